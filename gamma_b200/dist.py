"""Multi-GPU host logic: queries are sharded by rank, the index is replicated, and the per-rank top-k
(distances f32 + ids i64, [n_local, k]) are combined by ONE all-gather — the only exchange on the path
(SURVEY.md §8e).  Works with any torch.distributed backend: "nccl" on GPUs (bench.py, NVLink/NVSwitch),
"gloo" on CPU tensors (tests/test_dist_cpu.py)."""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """contiguous shard [lo, hi) of n queries for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_queries(xq, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(xq.shape[0], rank, world)
    return xq[lo:hi], lo, hi


def allgather_topk(D_local, I_local, n_total=None):
    """every rank gets the full [n_total, k] result in query order.  Equal shards use
    all_gather_into_tensor (one NCCL collective per tensor); ragged shards pad to the largest."""
    world = dist.get_world_size()
    k = D_local.shape[1]
    n_local = torch.tensor([D_local.shape[0]], device=D_local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    if all(s == mx for s in sizes):
        D_all = torch.empty(world * mx, k, dtype=D_local.dtype, device=D_local.device)
        I_all = torch.empty(world * mx, k, dtype=I_local.dtype, device=I_local.device)
        dist.all_gather_into_tensor(D_all, D_local.contiguous())
        dist.all_gather_into_tensor(I_all, I_local.contiguous())
        return D_all, I_all
    Dp = torch.zeros(mx, k, dtype=D_local.dtype, device=D_local.device)
    Ip = torch.full((mx, k), -1, dtype=I_local.dtype, device=I_local.device)
    Dp[:D_local.shape[0]] = D_local
    Ip[:I_local.shape[0]] = I_local
    Dl = [torch.empty_like(Dp) for _ in range(world)]
    Il = [torch.empty_like(Ip) for _ in range(world)]
    dist.all_gather(Dl, Dp)
    dist.all_gather(Il, Ip)
    D_all = torch.cat([d[:s] for d, s in zip(Dl, sizes)], 0)
    I_all = torch.cat([i[:s] for i, s in zip(Il, sizes)], 0)
    return D_all, I_all
