#!/usr/bin/env python
"""Supplementary measurement of the FLAT path (BASELINE.json configs[3]):
FLAT InnerProduct d=768, 5M vectors, batch=512, k=10 on one B200 — the tensor-core path
(tcgen05 3xTF32 chunked GEMM -> running candidate select -> exact fp32 re-score, flat_tc.cu).

Prints one JSON line in bench.py's shape.  The database is generated ON the device (a seeded torch mixture of
Gaussians, L2-normalised) and handed over with gb200_upload_raw_dev, so setup stays short; queries are fresh
samples of the same mixture.  Checks, inside the run: ids/distances of the first queries against an exact fp32
brute force (torch.matmul + topk, TF32 disabled) — the re-score makes the reported distances exact.

roofline: the GEMM kernel, bound "tensor"; achieved = EXECUTED tensor FLOP (3 x 2 n N d, the 3xTF32 split) per second
over the CUDA-event time of the whole search; peak = TF32 dense = measured bf16 peak / 2 (MEASURED_PEAKS.json).
"""
import argparse, ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=5_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", type=int, default=32, help="queries compared with the exact brute force")
    args = ap.parse_args()
    import torch
    from gamma_b200 import api
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(20240601)
    centres = torch.randn(4096, args.d, device=dev, generator=g)
    ix = api.B200FLAT(0)
    assert ix.Init(json.dumps({"metric_type": "InnerProduct"}), args.d) == 0
    t = time.time()
    chunk = 250_000

    def sample(n):
        a = torch.randint(0, 4096, (n,), device=dev, generator=g)
        x = centres[a] + 0.3 * torch.randn(n, args.d, device=dev, generator=g)
        return torch.nn.functional.normalize(x, dim=1).contiguous()

    keep = []  # database kept in torch as well for the exact check (first rows only when large)
    for s in range(0, args.N, chunk):
        x = sample(min(chunk, args.N - s))
        torch.cuda.synchronize()  # the library copies on its own stream: the rows must exist first
        ix.upload_raw_dev(x.data_ptr(), x.shape[0], first_vid=s)
        keep.append(x)
    xq = sample(args.batch)
    torch.cuda.synchronize()
    print("[flat] database %d x %d on device in %.1fs, %.2f GB" % (args.N, args.d, time.time() - t, ix.GetTotalMemBytes() / 1e9),
          file=sys.stderr, flush=True)
    n, k = args.batch, 10
    D = torch.empty(n, k, dtype=torch.float32, device=dev)
    I = torch.empty(n, k, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        rc = ix.search_dev(xq.data_ptr(), n, k, D.data_ptr(), I.data_ptr(), stream.cuda_stream, metric="InnerProduct")
        assert rc == 0, api.lib().gb200_last_error()

    step()
    torch.cuda.synchronize()
    # exact check
    nc = min(args.check, n)
    best_d = torch.full((nc, k), -float("inf"), device=dev)
    best_i = torch.full((nc, k), -1, dtype=torch.int64, device=dev)
    s0 = 0
    for x in keep:
        sc = xq[:nc] @ x.t()
        d_, i_ = sc.topk(k, dim=1)
        cd, ci = torch.cat([best_d, d_], 1), torch.cat([best_i, i_ + s0], 1)
        sel = cd.topk(k, dim=1)
        best_d, best_i = sel.values, torch.gather(ci, 1, sel.indices)
        s0 += x.shape[0]
    ids_ok = float((best_i == I[:nc]).float().mean())
    rel = float(((best_d - D[:nc]).abs() / best_d.abs().clamp(min=1e-6)).max())
    print("[flat] vs exact brute force on %d queries: ids identical %.4f, max rel distance error %.2e" % (nc, ids_ok, rel),
          file=sys.stderr, flush=True)
    del keep
    torch.cuda.empty_cache()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    l0 = ix.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = ix.launch_count() - l0
    # e2e through the host call
    xq_h = xq.cpu().numpy()
    sp = api._Base._sp("InnerProduct", -1, 0, 0, -api.FLT_MAX, api.FLT_MAX)
    D_h = np.empty((n, k), np.float32)
    I_h = np.empty((n, k), np.int64)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = api.lib().gb200_flat_search(ix.h, n, xq_h.ctypes.data, k, ctypes.byref(sp), None, 0, D_h.ctypes.data, I_h.ctypes.data)
        assert rc == 0
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    flop_exec = 3.0 * 2.0 * n * args.N * args.d
    out = dict(metric="QPS (FLAT InnerProduct d=%d, %d vecs, batch=%d, k=10)" % (args.d, args.N, n), value=n / (ms / 1e3), unit="queries/s",
               n_gpus=1, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, dtype="f32 (3xTF32 tensor-core candidates, exact fp32 re-score)",
               data="synthetic (device-generated mixture, L2-normalised)",
               config=dict(workload="FLAT IP d=%d N=%d batch=%d k=10, tensor-core path" % (args.d, args.N, n), l2="database (2 x %.1f GB) >> L2" % (args.N * args.d * 4 / 1e9)),
               roofline=dict(bound="tensor", achieved=flop_exec / (ms / 1e3) / 1e12, peak=tf32_peak, unit="TFLOP/s", frac=flop_exec / (ms / 1e3) / 1e12 / tf32_peak,
                             traffic=None, kernel="tc_gemm_tf32x3_kernel (share of the step: see the launch list)",
                             algorithmic_flop=2.0 * n * args.N * args.d, executed_tensor_flop=flop_exec,
                             hbm_algorithmic_bytes=args.N * args.d * 4, hbm_achieved_gbs=args.N * args.d * 4 / (ms / 1e3) / 1e9, hbm_peak=hbm_peak,
                             peak_source="TF32 dense = measured bf16 / 2 (MEASURED_PEAKS.json)" if peaks else "fallback"),
               e2e=dict(value=n / (e2e_ms / 1e3), unit="queries/s", h2d_bytes_per_step=n * args.d * 4, d2h_bytes_per_step=n * k * 12),
               gpu_launches=int(launches), check=dict(queries=nc, ids_identical_frac=ids_ok, max_rel_distance_err=rel))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
