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


def packed_topk_buffer(n_local, k, device):
    """one buffer for a rank's result: [n_local*k] f32 distances followed by [n_local*k] i64 ids.  Searching
    straight into its two views (bench.py does) makes the exchange a single collective with no packing copy."""
    buf = torch.empty(n_local * k * 12, dtype=torch.uint8, device=device)
    D = buf[: n_local * k * 4].view(torch.float32).view(n_local, k)
    I = buf[n_local * k * 4:].view(torch.int64).view(n_local, k)
    return buf, D, I


def unpack_topk(buf_all, world, n_local, k):
    per = n_local * k * 12
    blocks = buf_all.view(world, per)
    D = blocks[:, : n_local * k * 4].contiguous().view(torch.float32).view(world * n_local, k)
    I = blocks[:, n_local * k * 4:].contiguous().view(torch.int64).view(world * n_local, k)
    return D, I


def allgather_topk(D_local, I_local, equal_shards=None):
    """every rank gets the full [n_total, k] result in query order.  Equal shards: distances and ids travel in
    ONE all_gather_into_tensor of a packed byte buffer; ragged shards pad to the largest.  equal_shards=True
    skips the size exchange (the caller knows, e.g. a fixed batch per rank)."""
    world = dist.get_world_size()
    k = D_local.shape[1]
    if equal_shards is None:
        n_local = torch.tensor([D_local.shape[0]], device=D_local.device, dtype=torch.int64)
        sizes = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(sizes, n_local)
        sizes = [int(s.item()) for s in sizes]
    else:
        sizes = [D_local.shape[0]] * world
    mx = max(sizes)
    if all(s == mx for s in sizes):
        buf, D, I = packed_topk_buffer(mx, k, D_local.device)
        D.copy_(D_local)
        I.copy_(I_local)
        buf_all = torch.empty(world * buf.numel(), dtype=torch.uint8, device=D_local.device)
        dist.all_gather_into_tensor(buf_all, buf)
        return unpack_topk(buf_all, world, mx, k)
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
