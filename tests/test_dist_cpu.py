"""CPU, world_size 2, gloo: the multi-GPU host logic (query sharding + all-gather of per-shard top-k)
reproduces the single-process result.  The per-rank searcher here is the oracle restatement standing in
for the device library (the sharding/gather code is backend-agnostic; on GPUs the backend is NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_queries, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from gamma_b200 import dist as gdist
    from golden_util import Golden
    from oracle import gamma_oracle as go
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Golden("ivfpq_l2_d32_m8")
    xq = g["xq"][:n_queries]
    mine, lo, hi = gdist.shard_queries(xq, rank, world)
    D, I = go.ivfpq_search(mine, g["centroids"], g["pq"], g.lists, g["xb"], 10, g.nprobe, g.R, "L2", True,
                           keys=g["coarse_keys"][lo:hi], coarse_dis=g["coarse_dis"][lo:hi])
    D_all, I_all = gdist.allgather_topk(torch.from_numpy(D), torch.from_numpy(I))
    if rank == 0:
        np.save(os.path.join(out_dir, "D.npy"), D_all.numpy())
        np.save(os.path.join(out_dir, "I.npy"), I_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_queries", [12, 11])  # equal shards (all_gather_into_tensor) and ragged shards
def test_query_sharding_and_topk_allgather_world2(tmp_path, n_queries):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import Golden
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_queries, str(tmp_path)), nprocs=2, join=True)
    g = Golden("ivfpq_l2_d32_m8")
    D = np.load(tmp_path / "D.npy")
    I = np.load(tmp_path / "I.npy")
    assert np.array_equal(I, g["rank_I"][:n_queries]) and np.array_equal(D, g["rank_D"][:n_queries])


def test_shard_bounds_cover_everything():
    from gamma_b200 import dist as gdist
    for n in (0, 1, 7, 1024, 4097):
        for world in (1, 2, 3, 8):
            b = [gdist.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_packed_topk_buffer_roundtrip():
    """one buffer per rank ([n*k] f32 then [n*k] i64): what a rank searches into is what the all-gather ships"""
    import torch
    from gamma_b200 import dist as gdist
    world, n, k = 3, 5, 4
    bufs = []
    for r in range(world):
        buf, D, I = gdist.packed_topk_buffer(n, k, "cpu")
        D.copy_(torch.arange(n * k, dtype=torch.float32).view(n, k) + 100 * r)
        I.copy_(torch.arange(n * k, dtype=torch.int64).view(n, k) + 1000 * r)
        assert buf.numel() == n * k * 12 and D.data_ptr() == buf.data_ptr()
        bufs.append(buf)
    D_all, I_all = gdist.unpack_topk(torch.cat(bufs), world, n, k)
    assert D_all.shape == (world * n, k) and I_all.dtype == torch.int64
    for r in range(world):
        assert torch.equal(D_all[r * n:(r + 1) * n], torch.arange(n * k, dtype=torch.float32).view(n, k) + 100 * r)
        assert torch.equal(I_all[r * n:(r + 1) * n], torch.arange(n * k, dtype=torch.int64).view(n, k) + 1000 * r)
