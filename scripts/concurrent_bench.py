#!/usr/bin/env python
"""Concurrent Search through the host C-ABI (SURVEY §8b threading): T threads x batch B against one thread x batch T*B on
the headline index.  Prints one JSON line.  Uses bench.py's dataset/index helpers (and its GB200_BENCH_CACHE)."""
import ctypes, json, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from gamma_b200 import api
    w = bench.WORKLOADS["headline"]
    N, nlist, xb, xq_all = bench.build_dataset(w, 1.0)
    coarse, pq, list_no, codes = bench.build_index_state(w, N, nlist, xb, "cuda:0", (0, 1))
    ix = api.B200IVFPQ(0)
    assert ix.Init(json.dumps({"ncentroids": nlist, "nsubvector": w["M"], "metric_type": "L2", "nprobe": w["nprobe"]}), w["d"]) == 0
    ix.set_quantizers(coarse, pq)
    assert ix.append(list_no, np.arange(N, dtype=np.int64), codes) == 0
    for s in range(0, N, 1 << 21):
        ix.upload_raw(xb[s:s + (1 << 21)], first_vid=s)
    T, B, K = 8, 128, bench.K_TOP
    n = T * B
    xq = np.ascontiguousarray(xq_all[:n])
    sp = api._Base._sp("L2", w["nprobe"], bench.RECALL_NUM, True, -api.FLT_MAX, api.FLT_MAX)
    xq_pin = torch.from_numpy(xq).pin_memory()
    D = torch.empty(n, K, dtype=torch.float32).pin_memory()
    I = torch.empty(n, K, dtype=torch.int64).pin_memory()
    L = api.lib()

    def call(lo, hi):
        rc = L.gb200_ivfpq_search(ix.h, hi - lo, xq_pin.data_ptr() + lo * w["d"] * 4, K, ctypes.byref(sp), None, 0,
                                  D.data_ptr() + lo * K * 4, I.data_ptr() + lo * K * 8)
        assert rc == 0, L.gb200_last_error()

    reps = 30
    for _ in range(5):
        call(0, n)
    t0 = time.perf_counter()
    for _ in range(reps):
        call(0, n)
    serial = n * reps / (time.perf_counter() - t0)
    I_serial = I.numpy().copy()

    def worker(t):
        for _ in range(reps):
            call(t * B, (t + 1) * B)

    def run_threads():
        th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return n * reps / (time.perf_counter() - t0)

    # GB200_COALESCE* (capi.cu search_coalesced): 0 = every caller runs its own batch on its own context
    settings = [dict(), dict(GB200_COALESCE="0"), dict(GB200_COALESCE="1"), dict(GB200_COALESCE="2"), dict(GB200_COALESCE="4"),
                dict(GB200_COALESCE_WAIT_US="0"), dict(GB200_COALESCE_BALANCE="0")]
    table = []
    for st in settings:
        for kname in ("GB200_COALESCE", "GB200_COALESCE_WAIT_US", "GB200_COALESCE_BALANCE"):
            os.environ.pop(kname, None)
        os.environ.update(st)
        ix.reload_tuning()
        reps = 30
        run_threads()  # every search context of the pool allocates its workspaces here (cudaMalloc synchronises the device)
        reps = 200
        I.zero_()
        q = max(run_threads() for _ in range(2))
        table.append(dict(settings=st, qps=q, ids_identical=bool(np.array_equal(I.numpy(), I_serial))))
        print("[concurrent]", table[-1], file=sys.stderr, flush=True)
    for kname in ("GB200_COALESCE", "GB200_COALESCE_WAIT_US", "GB200_COALESCE_BALANCE"):
        os.environ.pop(kname, None)
    ix.reload_tuning()
    conc = table[0]["qps"]
    reps = 30
    same = bool(np.array_equal(I.numpy(), I_serial))
    t0 = time.perf_counter()
    for _ in range(reps):
        for t in range(T):
            call(t * B, (t + 1) * B)
    one_by_one = n * reps / (time.perf_counter() - t0)
    print(json.dumps(dict(workload=w["desc"], threads=T, batch_per_thread=B, qps_one_call_batch_1024=serial,
                          qps_8_threads_batch_128=conc, qps_1_thread_batch_128_back_to_back=one_by_one,
                          ids_identical=same and all(r["ids_identical"] for r in table), settings=table)), flush=True)


if __name__ == "__main__":
    main()
