"""Worker of tests/test_comm_gpu.py: one process per rank.  argv: rank world tmpdir device"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, tmp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    device = int(sys.argv[4]) if len(sys.argv) > 4 else rank
    import torch
    from gamma_b200 import api, builder, synth
    torch.cuda.set_device(device)
    dev = torch.device("cuda:%d" % device)
    N, d, nlist, M, n, k = 30000, 64, 64, 32, 96, 10
    xb = synth.mixture(N, d, 7, n_clusters=64)
    xq = synth.mixture(n * world, d, 8, n_clusters=64)
    cache = os.path.join(tmp, "ix.npz")
    if rank == 0:
        coarse, pq, list_no, codes = builder.build_ivfpq(xb, nlist, M, device=str(dev))
        np.savez(cache + ".tmp.npz", a=coarse, b=pq, c=list_no, d=codes)
        os.replace(cache + ".tmp.npz", cache)
    else:
        while not os.path.exists(cache):
            time.sleep(0.1)
        time.sleep(0.2)
        z = np.load(cache)
        coarse, pq, list_no, codes = z["a"], z["b"], z["c"], z["d"]
    ix = api.B200IVFPQ(device)
    assert ix.Init(json.dumps({"ncentroids": nlist, "nsubvector": M, "metric_type": "L2", "nprobe": 8}), d) == 0
    ix.set_quantizers(coarse, pq)
    assert ix.append(list_no, np.arange(N, dtype=np.int64), codes) == 0
    ix.upload_raw(xb)
    comm = api.Comm(device, rank, world, n * k * 12)
    with open(os.path.join(tmp, "h%d.tmp" % rank), "wb") as f:
        f.write(comm.handle_bytes())
    os.replace(os.path.join(tmp, "h%d.tmp" % rank), os.path.join(tmp, "h%d" % rank))
    handles = []
    for p in range(world):
        path = os.path.join(tmp, "h%d" % p)
        t0 = time.time()
        while not os.path.exists(path):
            assert time.time() - t0 < 120, "peer %d never published its handle" % p
            time.sleep(0.05)
        handles.append(open(path, "rb").read())
    comm.connect(handles)
    # the whole batch on this GPU alone = what the gathered result must equal
    rc, D_full, I_full = ix.Search(xq, k, nprobe=8, recall_num=50, metric="L2", has_rank=True)
    assert rc == 0
    xq_d = torch.from_numpy(np.ascontiguousarray(xq[rank * n:(rank + 1) * n])).to(dev)
    stream = torch.cuda.current_stream()
    ok = True
    for it in range(6):  # several epochs: both window buffers, flags running on
        base = comm.search_sharded(ix, xq_d.data_ptr(), n, k, stream.cuda_stream, nprobe=8, recall_num=50, metric="L2")
        torch.cuda.synchronize()
        host = torch.empty(world * comm.slot_bytes, dtype=torch.uint8)
        comm.read(host.data_ptr(), base, world * comm.slot_bytes, stream.cuda_stream, sync=True)
        for r in range(world):
            blk = host[r * comm.slot_bytes:(r + 1) * comm.slot_bytes]
            D = blk[:n * k * 4].view(torch.float32).numpy().reshape(n, k)
            I = blk[n * k * 4:n * k * 12].view(torch.int64).numpy().reshape(n, k)
            ok &= bool(np.array_equal(I, I_full[r * n:(r + 1) * n]) and np.array_equal(D, D_full[r * n:(r + 1) * n]))
    # pipelined form: call i hands back the gathered result of call i - 1; the batches alternate so that a stale or
    # early window would be noticed; mixed with the immediate form on the same comm
    xq2 = synth.mixture(n * world, d, 9, n_clusters=64)
    rc, D_full2, I_full2 = ix.Search(xq2, k, nprobe=8, recall_num=50, metric="L2", has_rank=True)
    assert rc == 0
    xq2_d = torch.from_numpy(np.ascontiguousarray(xq2[rank * n:(rank + 1) * n])).to(dev)
    expect = [(D_full, I_full), (D_full2, I_full2)]

    def check(base, which):
        host = torch.empty(world * comm.slot_bytes, dtype=torch.uint8)
        comm.read(host.data_ptr(), base, world * comm.slot_bytes, stream.cuda_stream, sync=True)
        good = True
        for r in range(world):
            blk = host[r * comm.slot_bytes:(r + 1) * comm.slot_bytes]
            D = blk[:n * k * 4].view(torch.float32).numpy().reshape(n, k)
            I = blk[n * k * 4:n * k * 12].view(torch.int64).numpy().reshape(n, k)
            good &= bool(np.array_equal(I, expect[which][1][r * n:(r + 1) * n]) and
                         np.array_equal(D, expect[which][0][r * n:(r + 1) * n]))
        return good

    prev = None
    for it in range(9):  # every window buffer more than twice
        which = it & 1
        base = comm.search_sharded(ix, (xq2_d if which else xq_d).data_ptr(), n, k, stream.cuda_stream, nprobe=8,
                                   recall_num=50, metric="L2", deferred=True)
        if it == 0:
            ok &= base is not None  # an immediate exchange ran before: its window is the "previous" one
        if base is not None and prev is not None:
            ok &= check(base, prev)
        prev = which
    ok &= check(comm.flush(stream.cuda_stream), prev)
    base = comm.search_sharded(ix, xq_d.data_ptr(), n, k, stream.cuda_stream, nprobe=8, recall_num=50, metric="L2")
    ok &= check(base, 0)
    st = comm.status()
    print(json.dumps(dict(rank=rank, ok=ok, status=st)), flush=True)
    comm.close()
    return 0 if (ok and st == 0) else 1


if __name__ == "__main__":
    sys.exit(main())
