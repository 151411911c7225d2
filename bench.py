#!/usr/bin/env python
"""bench.py — the hot path's headline metric on B200.

metric  : QPS @ recall@10 (d=128, 10M vecs, nprobe=32, batch=1024) — BASELINE.json
workload: IVFPQ d=128 nlist=16384 PQ32x8, 10M synthetic vectors, nprobe=32, recall_num=100,
          exact re-rank on, k=10, L2 (SURVEY.md §8d "headline shape").  `--workload c3` = PQ64x8 +
          range-filter bitmap + 1 % deletions, `--workload c2` = 1M/nlist 4096/nprobe 16/batch 256, `--workload c4` =
          FLAT inner product d=768 5M batch 512, `--workload c5` = 100M/nlist 65536/nprobe 64/batch 4096 built on the
          device (BASELINE.json configs).
step    : one Search of one batch (coarse quantiser + ADC scan + select + re-rank + top-k).
value   : device-timed (CUDA events), queries resident in HBM, L2 flushed between steps.
e2e     : the same Search through the public host C-ABI call with pinned HOST buffers
          (H2D of the queries and D2H of the results inside the timed region; N > 1: also the exchange and the D2H of
          the gathered result).
roofline: ADC scan kernel, algorithmic bytes = scanned postings x (code_size + 4)  (SURVEY §8d).
cpu_baseline / --impl reference: the reference's own CPU engine (oracle/_ref = unmodified
          GammaIVFPQIndex over faiss 1.7.1, compiled by oracle/Makefile) searching the SAME index on
          the host cores, on a bounded sample of the batch.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torch.distributed.run
(one rank per GPU, index replicated, queries sharded: every rank searches its own batch and its re-rank kernel stores
the top-k into every peer's result window over NVLink — gb200_ivfpq_search_sharded[_deferred], `--exchange`; weak
scaling, global batch = N x 1024; torch.distributed carries only the IPC handles and the timing reductions).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N, d, nlist, M, nprobe, batch, filter
    "headline": dict(N=10_000_000, d=128, nlist=16384, M=32, nprobe=32, batch=1024, filt=False,
                     desc="IVFPQ d=128 nlist=16384 PQ32x8 10M vecs nprobe=32 batch=1024 recall_num=100 rerank k=10 L2"),
    "c3": dict(N=10_000_000, d=128, nlist=16384, M=64, nprobe=32, batch=1024, filt=True,
               metric="QPS @ recall@10 (d=128, 10M vecs, PQ64x8, nprobe=32, batch=1024, range filter + deletions)",
               desc="IVFPQ d=128 nlist=16384 PQ64x8 10M vecs nprobe=32 batch=1024 + range-filter bitmap (30% pass) + 1% deleted"),
    "c2": dict(N=1_000_000, d=128, nlist=4096, M=32, nprobe=16, batch=256, filt=False,
               metric="QPS @ recall@10 (d=128, 1M vecs, nlist=4096, nprobe=16, batch=256)",
               desc="IVFPQ d=128 nlist=4096 PQ32x8 1M vecs nprobe=16 batch=256"),
    # BASELINE.json configs[4]: built ON the device (seeded torch mixture generated chunk by chunk, trained with the setup
    # tooling, encoded + appended by gb200_ivfpq_add_stored) — 51 GB of raw vectors never touch the host
    "c5": dict(N=100_000_000, d=128, nlist=65536, M=32, nprobe=64, batch=4096, filt=False, device_build=True,
               cpu_N=10_000_000, metric="QPS @ recall@10 (d=128, 100M vecs, nlist=65536, nprobe=64, batch=4096)",
               desc="IVFPQ d=128 nlist=65536 PQ32x8 100M vecs nprobe=64 batch=4096 recall_num=100 rerank k=10 L2"),
    # BASELINE.json configs[3]: the tensor-core flat path
    "c4": dict(kind="flat", N=5_000_000, d=768, batch=512, metric="InnerProduct",
               desc="FLAT InnerProduct d=768 5M vecs batch=512 k=10 (tcgen05 3xTF32 candidates, exact fp32 re-score)"),
}
K_TOP, RECALL_NUM = 10, 100
METRIC = "QPS @ recall@10 (d=128, 10M vecs, nprobe=32, batch=1024)"


class DeviceData:
    """Seeded mixture of Gaussians generated ON the device, chunk by chunk; chunk [s, e) always comes out the same (its
    generator is seeded from s), so the rows can be regenerated for the ground truth and for the CPU engine's subset."""
    CHUNK = 1_000_000

    def __init__(self, d, dev, normalize=False, n_clusters=4096, seed=20240601, spread=0.3):
        import torch
        self.d, self.dev, self.normalize, self.seed, self.spread = d, dev, normalize, seed, spread
        g = torch.Generator(device=dev).manual_seed(seed + 7919)
        self.centres = torch.randn(n_clusters, d, device=dev, generator=g)

    def rows(self, s, e, salt=0):
        import torch
        g = torch.Generator(device=self.dev).manual_seed(self.seed + 1000003 * salt + s)
        a = torch.randint(0, self.centres.shape[0], (e - s,), device=self.dev, generator=g)
        x = self.centres[a] + self.spread * torch.randn(e - s, self.d, device=self.dev, generator=g)
        if self.normalize:
            x = torch.nn.functional.normalize(x, dim=1)
        return x.contiguous()

    def queries(self, n, rank):
        return self.rows(0, n, salt=17 + rank)


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs (B200_PROFILING.md).  NVML from a thread every
    few milliseconds (the timed region of this bench is ~10 ms, far below nvidia-smi's loop period); the nvidia-smi
    recipe line is the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    MASKS = dict(sw_power_cap=0x4, hw_slowdown=0x8, sw_thermal_slowdown=0x20, hw_thermal_slowdown=0x40)

    def __init__(self, gpu_index, uuid=None):
        self.gpu = gpu_index
        self.uuid = uuid
        self.lines = []
        self.samples = []  # (sm_mhz, reasons bitmask)
        self.proc = None
        self.h = None
        self.max_mhz = None
        self.stop_flag = False
        self.th = None
        self.how = None

    def _nvml_start(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        if self.uuid:
            try:
                u = str(self.uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(u if u.startswith("GPU-") else "GPU-" + u)
            except Exception:
                h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")

        def loop():
            while not self.stop_flag:
                try:
                    self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(get_reasons(h))))
                except Exception:
                    pass
                time.sleep(0.003)

        loop_once = (float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(get_reasons(h)))  # fails here, not in the thread
        del loop_once
        self.th = threading.Thread(target=loop, daemon=True)
        self.th.start()
        self.how = "nvml"

    def start(self):
        try:
            self._nvml_start()
            return
        except Exception:
            self.th = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            self.how = "nvidia-smi"
        except Exception:
            self.proc = None

    def mark(self):
        """number of samples so far (the caller brackets the timed region with two marks)"""
        return len(self.samples) if self.th else len(self.lines)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, m0=None, m1=None):
        if self.th:
            self.stop_flag = True
            self.th.join(timeout=1.0)
            sm = [s[0] for s in self.samples]
            bits = 0
            for s in self.samples:
                bits |= s[1]
            reasons = sorted(k for k, v in self.MASKS.items() if bits & v)
            out = dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=self.max_mhz, reasons=reasons,
                       samples=len(sm), how="nvml thread, 3 ms period, over pre-load + timed region + post-load")
            if m0 is not None and m1 is not None:
                tr = sm[m0:m1]
                out["samples_in_timed_region"] = len(tr)
                if tr:
                    out["sm_mhz_timed_region"] = float(np.median(tr))
            return out
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock sampling unavailable"], samples=0)
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), how="nvidia-smi -lms 100 over pre-load + timed region + post-load")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _cache_path(tag):
    """dev only: GB200_BENCH_CACHE=<dir> keeps the seeded synthetic set and the trained/encoded index between
    bench.py invocations on the same box (setup is outside every timed region; the data are identical)."""
    c = os.environ.get("GB200_BENCH_CACHE")
    if not c:
        return None
    os.makedirs(c, exist_ok=True)
    return os.path.join(c, tag)


def build_dataset(w, scale):
    from gamma_b200 import synth
    N = max(int(w["N"] * scale), 50 * 256)
    nlist = w["nlist"] if scale == 1.0 else max(64, int(w["nlist"] * scale))
    t = time.time()
    cache = _cache_path("data_%d_%d" % (N, w["d"]))
    if cache and os.path.exists(cache + ".xb.npy"):
        xb = np.load(cache + ".xb.npy")
    else:
        xb = synth.base_vectors(N, w["d"])
        if cache and int(os.environ.get("RANK", "0")) == 0:
            np.save(cache + ".xb.npy", xb)
    xq_all = synth.query_vectors(w["batch"] * 8, w["d"])
    log("synthetic data N=%d d=%d in %.1fs" % (N, w["d"], time.time() - t))
    return N, nlist, xb, xq_all


def build_index_state(w, N, nlist, xb, device, dist_ctx):
    """rank 0 trains + encodes on its GPU (torch, setup only), other ranks receive the result."""
    import torch
    from gamma_b200 import builder
    rank, world = dist_ctx
    M, d = w["M"], w["d"]
    if rank == 0:
        t = time.time()
        cache = _cache_path("index_%d_%d_%d_%d.npz" % (N, d, nlist, M))
        if cache and os.path.exists(cache):
            z = np.load(cache)
            coarse, pq, list_no, codes = z["coarse"], z["pq"], z["list_no"], z["codes"]
        else:
            coarse, pq, list_no, codes = builder.build_ivfpq(xb, nlist, M, device=device)
            if cache:
                np.savez(cache, coarse=coarse, pq=pq, list_no=list_no, codes=codes)
        log("trained + encoded in %.1fs" % (time.time() - t))
    if world > 1:
        import torch.distributed as dist
        dev = torch.device(device)
        tc = torch.from_numpy(coarse).to(dev) if rank == 0 else torch.empty(nlist, d, device=dev)
        tp = torch.from_numpy(pq).to(dev) if rank == 0 else torch.empty(M, 256, d // M, device=dev)
        tl = torch.from_numpy(list_no).to(dev) if rank == 0 else torch.empty(N, dtype=torch.int32, device=dev)
        tcd = torch.from_numpy(codes).to(dev) if rank == 0 else torch.empty(N, M, dtype=torch.uint8, device=dev)
        for t_ in (tc, tp, tl, tcd):
            dist.broadcast(t_, 0)
        coarse, pq, list_no, codes = tc.cpu().numpy(), tp.cpu().numpy(), tl.cpu().numpy(), tcd.cpu().numpy()
    return coarse, pq, list_no, codes


def make_filter(w, N):
    from gamma_b200 import synth
    if not w["filt"]:
        return [], None
    flags = (synth.filter_field(N) < 30).astype(np.uint8)
    dele = synth.deleted_docs(N, 0.01)
    return [(0, N - 1, False, flags)], dele


def ground_truth(xb_t, xq_t, k, valid_mask_t=None, chunk=1 << 20):
    import torch
    n = xq_t.shape[0]
    best_d = torch.full((n, k), float("inf"), device=xq_t.device)
    best_i = torch.full((n, k), -1, dtype=torch.int64, device=xq_t.device)
    qn = (xq_t * xq_t).sum(1, keepdim=True)
    for s in range(0, xb_t.shape[0], chunk):
        xs = xb_t[s:s + chunk]
        dist = qn + (xs * xs).sum(1)[None, :] - 2.0 * (xq_t @ xs.t())
        if valid_mask_t is not None:
            dist = dist.masked_fill(~valid_mask_t[s:s + chunk][None, :], float("inf"))
        d, i = dist.topk(k, dim=1, largest=False)
        cat_d = torch.cat([best_d, d], 1)
        cat_i = torch.cat([best_i, i + s], 1)
        sel = cat_d.topk(k, dim=1, largest=False)
        best_d, best_i = sel.values, torch.gather(cat_i, 1, sel.indices)
    return best_i


def recall_at_k(I, gt):
    hit = 0
    for a, b in zip(I, gt):
        hit += len(set(int(x) for x in a if x >= 0) & set(int(x) for x in b))
    return hit / float(gt.shape[0] * gt.shape[1])


def build_reference(w, N, nlist, xb, coarse, pq, list_no, codes, dele):
    from oracle import ref
    t = time.time()
    model_json = json.dumps({"ncentroids": nlist, "nsubvector": w["M"], "metric_type": "L2", "nprobe": w["nprobe"]})
    wd = tempfile.mkdtemp(prefix="oref_bench_")
    r = ref.RefIndex(w["d"], "IVFPQ", model_json, indexing_size=N, bitmap_bits=max(2 * N, 1024), work_dir=wd)
    r._own_dir = True
    for s in range(0, N, 1 << 20):
        r.add_raw(xb[s:s + (1 << 20)])
    r.set_trained(coarse, pq)
    r.inject_postings(list_no, np.arange(N, dtype=np.int64), codes)
    if dele is not None:
        for doc in dele:
            r.delete(int(doc))
    log("reference CPU engine loaded with the same index in %.1fs" % (time.time() - t))
    return r


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def time_reference(r, w, xq, filters, n_queries, reps, warm=1):
    from oracle import ref
    ref.set_threads(host_cores())
    cores = ref.max_threads()
    rj = json.dumps({"nprobe": w["nprobe"], "recall_num": RECALL_NUM, "metric_type": "L2", "parallel_on_queries": 1})
    q = xq[:n_queries]
    times = []
    I = None
    for it in range(warm + reps):
        t = time.perf_counter()
        D, I = r.search(q, K_TOP, rj, has_rank=True, filters=filters)
        dt = time.perf_counter() - t
        if it >= warm:
            times.append(dt)
    return n_queries / float(np.median(times)), cores, times, I


def build_on_device(w, local, rank, with_lib, scale=1.0):
    """Workloads too large to stage on the host (c5): rows are generated on the device chunk by chunk (DeviceData),
    the quantizers are trained with the setup tooling on the first rows, and — with_lib — every chunk goes
    gb200_upload_raw_dev -> gb200_ivfpq_add_stored (device encode + append).  Returns the index (or None), the
    trained state, and host copies of the first cpu_N rows with their list numbers / codes for the CPU engine."""
    import torch
    from gamma_b200 import builder
    dev = torch.device("cuda:%d" % local)
    N, d, nlist, M, cpu_N = w["N"], w["d"], w["nlist"], w["M"], w["cpu_N"]
    if scale != 1.0:  # dev only
        N, nlist = max(int(N * scale), 200_000), max(256, int(nlist * scale))
        cpu_N = min(cpu_N, N)
    data = DeviceData(d, dev)
    t = time.time()
    train_n = min(N, 39 * nlist)
    xt = torch.cat([data.rows(s, min(s + DeviceData.CHUNK, train_n)) for s in range(0, train_n, DeviceData.CHUNK)])
    coarse_t = builder.kmeans(xt, nlist, 10, 1234)
    a = builder._assign(xt, coarse_t, (coarse_t * coarse_t).sum(1))
    sub = slice(0, min(train_n, 400_000))  # 256 points per PQ centroid are plenty (faiss caps the training set likewise)
    pq_t = builder.train_pq((xt[sub] - coarse_t[a[sub]]).contiguous(), M, 25, 1235)
    del xt, a
    torch.cuda.empty_cache()
    coarse, pq = coarse_t.cpu().numpy(), pq_t.cpu().numpy()
    log("trained %d coarse centroids + PQ%dx8 on %d rows in %.1fs" % (nlist, M, train_n, time.time() - t))
    ix = None
    if with_lib:
        from gamma_b200 import api
        ix = api.B200IVFPQ(local)
        model_json = json.dumps({"ncentroids": nlist, "nsubvector": M, "metric_type": "L2", "nprobe": w["nprobe"]})
        if ix.Init(model_json, d) != 0:
            raise SystemExit("gb200 create failed: %s" % api.lib().gb200_last_error().decode())
        ix.set_quantizers(coarse, pq)
    want_cpu = rank == 0
    xb_cpu = np.empty((cpu_N, d), np.float32) if want_cpu else None
    ln_cpu = np.empty(cpu_N, np.int32) if want_cpu else None
    cd_cpu = np.empty((cpu_N, M), np.uint8) if want_cpu else None
    t = time.time()
    for s in range(0, N, DeviceData.CHUNK):
        e = min(N, s + DeviceData.CHUNK)
        if not with_lib and s >= cpu_N:
            break
        x = data.rows(s, e)
        torch.cuda.synchronize()  # the library copies on its own stream: the rows must exist first
        keep = want_cpu and s < cpu_N
        if with_lib:
            ix.upload_raw_dev(x.data_ptr(), e - s, first_vid=s)
            got = ix.add_stored(s, e - s, want_codes=keep)
        else:
            ln_t, cd_t = builder.encode(x, coarse_t, pq_t)
            got = (ln_t.cpu().numpy(), cd_t.cpu().numpy())
        if keep:
            m = min(e, cpu_N) - s
            xb_cpu[s:s + m] = x[:m].cpu().numpy()
            ln_cpu[s:s + m] = got[0][:m]
            cd_cpu[s:s + m] = got[1][:m]
        del x
    log("%s %d rows in %.1fs" % ("uploaded + device-encoded + appended" if with_lib else "encoded the CPU subset of", N, time.time() - t))
    return dict(ix=ix, data=data, coarse=coarse, pq=pq, xb_cpu=xb_cpu, ln_cpu=ln_cpu, cd_cpu=cd_cpu, N=N, nlist=nlist, cpu_N=cpu_N)


def ground_truth_regenerated(data, N, xq_t, k):
    """exact top-k over rows regenerated chunk by chunk (the database exists only inside the library)"""
    import torch
    n = xq_t.shape[0]
    best_d = torch.full((n, k), float("inf"), device=xq_t.device)
    best_i = torch.full((n, k), -1, dtype=torch.int64, device=xq_t.device)
    qn = (xq_t * xq_t).sum(1, keepdim=True)
    for s in range(0, N, DeviceData.CHUNK):
        xs = data.rows(s, min(N, s + DeviceData.CHUNK))
        dist = qn + (xs * xs).sum(1)[None, :] - 2.0 * (xq_t @ xs.t())
        d_, i_ = dist.topk(k, dim=1, largest=False)
        cat_d, cat_i = torch.cat([best_d, d_], 1), torch.cat([best_i, i_ + s], 1)
        sel = cat_d.topk(k, dim=1, largest=False)
        best_d, best_i = sel.values, torch.gather(cat_i, 1, sel.indices)
    return best_i


def main_flat(args, w, rank, world, local, device):
    """--workload c4: FLAT InnerProduct d=768, 5M vectors, batch 512, k=10 — the tensor-core flat path (tcgen05 3xTF32
    chunked GEMM -> running candidate select -> exact fp32 re-score).  Database generated on the device and replicated
    per rank; queries sharded by rank (weak scaling), one NCCL all-gather of the packed top-k for N > 1.
    roofline: bound "tensor" for the GEMM kernel, achieved = ALGORITHMIC flops 2 n N d over the CUDA-event
    time of the search; peak = TF32 dense = measured bf16 / 2."""
    import torch
    N, d, n, metric = w["N"], w["d"], w["batch"], w["metric"]
    if args.scale != 1.0:
        N = max(int(N * args.scale), 20000)
    ip = metric != "L2"
    k = K_TOP
    fj = json.dumps({"metric_type": metric, "parallel_on_queries": 0})
    flat_metric = "QPS (FLAT %s d=%d, %d vecs, batch=%d, k=%d)" % (metric, d, N, n, k)
    # ------------------------------------------------------------------ reference arm: GammaFLATIndex on the host cores
    if args.impl == "reference":
        from oracle import ref
        data = DeviceData(d, torch.device(device), normalize=ip) if torch.cuda.is_available() else None
        if data is None:
            raise SystemExit("bench.py --workload c4 generates its rows on the device: needs a GPU")
        t = time.time()
        r = ref.RefIndex(d, "FLAT", json.dumps({"metric_type": metric}), indexing_size=N, bitmap_bits=max(2 * N, 1024))
        for s in range(0, N, DeviceData.CHUNK):
            r.add_raw(data.rows(s, min(N, s + DeviceData.CHUNK)).cpu().numpy())
        xq = data.queries(n, 0).cpu().numpy()
        log("reference FLAT engine loaded with %d rows in %.1fs" % (N, time.time() - t))
        ref.set_threads(host_cores())
        nq = min(args.cpu_queries, 16, n)
        ts = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            r.search(xq[:nq], k, fj)
            if it >= args.warmup:
                ts.append(time.perf_counter() - t0)
        ms = 1e3 * float(np.mean(ts))
        val = nq / (ms / 1e3)
        out = dict(metric=flat_metric, value=val, unit="queries/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                   ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                   impl="reference", config=dict(workload=w["desc"], N=N, queries_per_step=nq),
                   cpu_baseline=dict(value=val, unit="queries/s", cores=ref.max_threads(), kind="reference",
                                     sample="%d of the %d-query batch per step, %d steps, parallel_on_queries=0 "
                                            "(OpenMP over the database, gamma_index_flat.cc:250-291)" % (nq, n, args.steps)),
                   e2e=dict(value=val, unit="queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(out), flush=True)
        return 0
    # ------------------------------------------------------------------ our arm
    from gamma_b200 import api
    from gamma_b200 import dist as gdist
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device(device)
    data = DeviceData(d, dev, normalize=ip)
    ix = api.B200FLAT(local)
    if ix.Init(json.dumps({"metric_type": metric}), d) != 0:
        raise SystemExit("gb200 flat create failed: %s" % api.lib().gb200_last_error().decode())
    t = time.time()
    for s in range(0, N, DeviceData.CHUNK):
        x = data.rows(s, min(N, s + DeviceData.CHUNK))
        torch.cuda.synchronize()
        ix.upload_raw_dev(x.data_ptr(), x.shape[0], first_vid=s)
    xq_d = data.queries(n, rank)
    torch.cuda.synchronize()
    log("database %d x %d on the device in %.1fs, %.2f GB" % (N, d, time.time() - t, ix.GetTotalMemBytes() / 1e9))
    out_d, D_d, I_d = gdist.packed_topk_buffer(n, k, dev)
    out_bytes = out_d.numel()
    stream = torch.cuda.current_stream()
    if world > 1:
        import torch.distributed as dist
        out_all = torch.empty(world * out_bytes, dtype=torch.uint8, device=dev)

    def step_dev():
        rc_ = ix.search_dev(xq_d.data_ptr(), n, k, D_d.data_ptr(), I_d.data_ptr(), stream.cuda_stream, metric=metric)
        assert rc_ == 0, api.lib().gb200_last_error()
        if world > 1:
            dist.all_gather_into_tensor(out_all, out_d)

    step_dev()
    torch.cuda.synchronize()
    I_ours, D_ours = I_d.cpu().numpy().copy(), D_d.cpu().numpy().copy()
    # exact check against an fp32 brute force over regenerated rows (first 32 queries)
    nc = min(32, n)
    best_d = torch.full((nc, k), -float("inf") if ip else float("inf"), device=dev)
    best_i = torch.full((nc, k), -1, dtype=torch.int64, device=dev)
    for s in range(0, N, DeviceData.CHUNK):
        xs = data.rows(s, min(N, s + DeviceData.CHUNK))
        sc = xq_d[:nc] @ xs.t()
        if not ip:
            sc = (xq_d[:nc] * xq_d[:nc]).sum(1, keepdim=True) + (xs * xs).sum(1)[None, :] - 2.0 * sc
        d_, i_ = sc.topk(k, dim=1, largest=ip)
        cd, ci = torch.cat([best_d, d_], 1), torch.cat([best_i, i_ + s], 1)
        sel = cd.topk(k, dim=1, largest=ip)
        best_d, best_i = sel.values, torch.gather(ci, 1, sel.indices)
    ids_ok = float((best_i.cpu().numpy() == I_ours[:nc]).mean())
    rel = float(((best_d - D_d[:nc]).abs() / best_d.abs().clamp(min=1e-6)).max())
    log("vs exact fp32 brute force on %d queries: ids identical %.4f, max rel distance error %.2e" % (nc, ids_ok, rel))
    torch.cuda.empty_cache()

    ix.set_profiling(True)
    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local, gpu_uuid)
    sampler.start()
    for _ in range(3):
        step_dev()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = ix.launch_count()
    mark0 = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()  # the database (15 GB, twice with its TF32 companion) is far larger than L2: no flush needed
    e1.record(stream)
    torch.cuda.synchronize()
    mark1 = sampler.mark()
    launches = ix.launch_count() - launches0 + (args.steps if world > 1 else 0)
    if world > 1:
        dist.barrier()
    for _ in range(3):
        step_dev()
    torch.cuda.synchronize()
    clocks = sampler.stop(mark0, mark1)
    tm = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item()) / args.steps
    value = world * n / (ms / 1e3)

    # ---- e2e through the host C-ABI call (pinned host buffers, H2D + D2H inside); N > 1: + exchange + D2H of the gather
    xq_pin = xq_d.cpu().pin_memory()
    D_pin = torch.empty(n, k, dtype=torch.float32).pin_memory()
    I_pin = torch.empty(n, k, dtype=torch.int64).pin_memory()
    sp = api._Base._sp(metric, -1, 0, 0, -api.FLT_MAX, api.FLT_MAX)
    if world > 1:
        out_all_pin = torch.empty(world * out_bytes, dtype=torch.uint8).pin_memory()
        xq_stage = torch.empty_like(xq_d)

    def step_host():
        if world == 1:
            rc_ = api.lib().gb200_flat_search(ix.h, n, xq_pin.data_ptr(), k, ctypes.byref(sp), None, 0, D_pin.data_ptr(),
                                              I_pin.data_ptr())
            assert rc_ == 0, api.lib().gb200_last_error()
            return
        xq_stage.copy_(xq_pin, non_blocking=True)
        rc_ = ix.search_dev(xq_stage.data_ptr(), n, k, D_d.data_ptr(), I_d.data_ptr(), stream.cuda_stream, metric=metric)
        assert rc_ == 0, api.lib().gb200_last_error()
        dist.all_gather_into_tensor(out_all, out_d)
        out_all_pin.copy_(out_all, non_blocking=True)
        stream.synchronize()

    for _ in range(2):
        step_host()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_qps = world * n * args.steps / float(te.item())
    if world == 1:
        assert np.array_equal(I_pin.numpy(), I_ours), "host-API result differs from device-API result"
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf32_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
    hbm_peak, _ = measured_peaks()
    flop_alg = 2.0 * n * N * d
    roofline = dict(bound="tensor", achieved=flop_alg / (ms / 1e3) / 1e12, peak=tf32_peak, unit="TFLOP/s",
                    frac=flop_alg / (ms / 1e3) / 1e12 / tf32_peak, traffic=None,
                    kernel="tc_gemm_tf32x3_kernel (one launch per database chunk; time = CUDA events around the whole search)",
                    algorithmic_flop_per_step=flop_alg, executed_tensor_flop_per_step=3.0 * flop_alg,
                    hbm_algorithmic_bytes=float(N) * d * 4, hbm_achieved_gbs=float(N) * d * 4 / (ms / 1e3) / 1e9, hbm_peak=hbm_peak,
                    peak_source="TF32 dense = measured bf16 / 2 (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)")
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            from oracle import ref
            t = time.time()
            r = ref.RefIndex(d, "FLAT", json.dumps({"metric_type": metric}), indexing_size=N, bitmap_bits=max(2 * N, 1024))
            for s in range(0, N, DeviceData.CHUNK):
                r.add_raw(data.rows(s, min(N, s + DeviceData.CHUNK)).cpu().numpy())
            log("reference FLAT engine loaded with the same %d rows in %.1fs" % (N, time.time() - t))
            ref.set_threads(host_cores())
            nq = min(args.cpu_queries, 16, n)
            xq_h = xq_d.cpu().numpy()
            ts = []
            I_cpu = None
            for it in range(3):
                t0 = time.perf_counter()
                D_cpu, I_cpu = r.search(xq_h[:nq], k, fj)
                if it >= 1:
                    ts.append(time.perf_counter() - t0)
            cpu = dict(value=nq / float(np.median(ts)), unit="queries/s", cores=ref.max_threads(), kind="reference",
                       sample="first %d queries of the batch over all %d rows, median of 2 runs after 1 warm-up, "
                              "parallel_on_queries=0" % (nq, N),
                       ids_identical_frac=float((I_cpu == I_ours[:nq]).mean()),
                       distances_identical_frac=float((D_cpu == D_ours[:nq]).mean()))
            r.close()
        except Exception as e:
            cpu = dict(value=None, unit="queries/s", cores=None, kind="reference", sample="failed: %r" % (e,))
    out = dict(metric=flat_metric, value=value, unit="queries/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
               ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f32 (3xTF32 tensor-core candidates, exact fp32 re-score)", data="synthetic (device-generated mixture)",
               config=dict(workload=w["desc"], N=N, d=d, batch_per_gpu=n, global_batch=world * n, k=k,
                           parallelism="query-sharded x%d, database replicated, NCCL all-gather of top-k" % world,
                           l2="database (2 x %.1f GB with its TF32 companion) >> L2: no flush" % (N * d * 4 / 1e9),
                           scaled_down=args.scale != 1.0),
               roofline=roofline, cpu_baseline=cpu,
               e2e=dict(value=e2e_qps, unit="queries/s", h2d_bytes_per_step=int(n * d * 4),
                        d2h_bytes_per_step=int(n * k * 12 * (world if world > 1 else 1))),
               gpu_launches=int(launches), clocks=clocks,
               check=dict(queries=nc, ids_identical_frac=ids_ok, max_rel_distance_err=rel))
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="dev only: shrink N and nlist (result not valid)")
    ap.add_argument("--cpu-queries", type=int, default=256, help="bounded sample of the batch for the CPU engine")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variants", default="", help="dev only: ';'-separated ENV=VAL[,ENV=VAL] sets to time after the main run")
    ap.add_argument("--exchange", default="p2p-deferred", choices=["p2p", "p2p-deferred", "nccl"],
                    help="N > 1: result exchange by peer stores over NVLink behind the C-ABI — p2p-deferred (default): a "
                         "step waits for the peers' results of the step before it (gb200_ivfpq_search_sharded_deferred), "
                         "p2p: for this step's (gb200_ivfpq_search_sharded) — or by an NCCL all-gather of the packed top-k")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return 0  # rank 0 alone runs and prints the reference arm

    import torch
    have_gpu = torch.cuda.is_available()
    if args.impl == "ours" and not have_gpu:
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    device = "cuda:%d" % local if have_gpu else "cpu"
    if have_gpu:
        torch.cuda.set_device(local)
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))

    if w.get("kind") == "flat":
        return main_flat(args, w, rank, world, local, device)
    devb = None
    n = w["batch"]
    if w.get("device_build"):
        if not have_gpu:
            raise SystemExit("bench.py --workload %s builds its index on the device: needs a GPU" % args.workload)
        devb = build_on_device(w, local, rank, with_lib=args.impl == "ours", scale=args.scale)
        N, nlist = devb["N"], devb["nlist"]
        coarse, pq = devb["coarse"], devb["pq"]
        # the CPU engine gets the first cpu_N rows (stated in the JSON line); ours searches all N
        xb, list_no, codes = devb["xb_cpu"], devb["ln_cpu"], devb["cd_cpu"]
        filters, dele = [], None
        xq = devb["data"].queries(n, rank).cpu().numpy()
    else:
        N, nlist, xb, xq_all = build_dataset(w, args.scale)
        dist_ctx = (rank, world) if args.impl == "ours" else (0, 1)
        coarse, pq, list_no, codes = build_index_state(w, N, nlist, xb, device, dist_ctx)
        filters, dele = make_filter(w, N)
        # every rank gets its own batch of fresh queries (weak scaling); rank r uses slice r
        xq = np.ascontiguousarray(xq_all[(rank % 8) * n:(rank % 8 + 1) * n])
    N_cpu = devb["cpu_N"] if devb else N
    cpu_note = (" on the first %d of the %d vectors (same quantizers, lists %.0fx shorter)" % (N_cpu, N, N / N_cpu)) if devb else ""

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        r = build_reference(w, N_cpu, nlist, xb, coarse, pq, list_no, codes, dele)
        nq = min(args.cpu_queries, n)
        qps_list = []
        from oracle import ref
        # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the reference engine gets every core this
        # process may run on, whatever the launcher put in the environment
        ref.set_threads(host_cores())
        cores = ref.max_threads()
        rj = json.dumps({"nprobe": w["nprobe"], "recall_num": RECALL_NUM, "metric_type": "L2", "parallel_on_queries": 1})
        for it in range(args.warmup + args.steps):
            t = time.perf_counter()
            r.search(xq[:nq], K_TOP, rj, has_rank=True, filters=filters)
            dt = time.perf_counter() - t
            if it >= args.warmup:
                qps_list.append(dt)
        ms = 1e3 * float(np.mean(qps_list))
        val = nq / (ms / 1e3)
        out = dict(metric=w.get("metric", METRIC), value=val, unit="queries/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                   ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                   data="synthetic", impl="reference",
                   config=dict(workload=w["desc"], N=N, nlist=nlist, queries_per_step=nq, scaled_down=args.scale != 1.0 or bool(devb),
                               cpu_N=N_cpu),
                   cpu_baseline=dict(value=val, unit="queries/s", cores=cores, kind="reference",
                                     sample="%d of the %d-query batch per step, %d steps, OpenMP over queries%s" % (nq, n, args.steps, cpu_note)),
                   e2e=dict(value=val, unit="queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(out), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm
    from gamma_b200 import api
    t = time.time()
    if devb:
        ix = devb["ix"]
    else:
        model_json = json.dumps({"ncentroids": nlist, "nsubvector": w["M"], "metric_type": "L2", "nprobe": w["nprobe"]})
        ix = api.B200IVFPQ(local)
        rc = ix.Init(model_json, w["d"])
        if rc != 0:
            raise SystemExit("gb200 create failed: %s" % api.lib().gb200_last_error().decode())
        ix.set_quantizers(coarse, pq)
        rc = ix.append(list_no, np.arange(N, dtype=np.int64), codes)
        assert rc == 0, api.lib().gb200_last_error()
        for s in range(0, N, 1 << 21):
            ix.upload_raw(xb[s:s + (1 << 21)], first_vid=s)
        if dele is not None:
            ix.set_deleted(dele, True)
        if filters:
            ix.set_filters(filters)
    log("device mirror built in %.1fs, %.2f GB" % (time.time() - t, ix.GetTotalMemBytes() / 1e9))

    dev = torch.device(device)
    xq_d = torch.from_numpy(xq).to(dev)
    # distances and ids of one rank live in ONE buffer ([n*k] f32 followed by [n*k] i64) so that the multi-GPU
    # exchange is a single all-gather (two small collectives cost ~35 us of launch latency each)
    from gamma_b200 import dist as gdist
    out_d, D_d, I_d = gdist.packed_topk_buffer(n, K_TOP, dev)
    out_bytes = out_d.numel()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()
    comm = None
    p2p = world > 1 and args.exchange in ("p2p", "p2p-deferred")
    deferred = p2p and args.exchange == "p2p-deferred"  # a step waits for the peers' results of the step before it
    if world > 1:
        import torch.distributed as dist
        out_all = torch.empty(world * out_bytes, dtype=torch.uint8, device=dev)  # rank r's block at r * out_bytes
    if p2p:
        # the exchange lives behind the C-ABI: every rank stores its top-k into every peer's window over NVLink; the
        # host only carries the opaque IPC handles between the processes once
        comm = api.Comm(local, rank, world, out_bytes)
        hs = [None] * world
        dist.all_gather_object(hs, comm.handle_bytes())
        comm.connect(hs)
    gathered = [0]  # device address of the gathered window of the last sharded search

    def step_plain():
        rc_ = ix.search_dev(xq_d.data_ptr(), n, K_TOP, D_d.data_ptr(), I_d.data_ptr(), stream.cuda_stream,
                            nprobe=w["nprobe"], recall_num=RECALL_NUM, metric="L2", has_rank=True)
        assert rc_ == 0, api.lib().gb200_last_error()

    def step_dev():
        if p2p:
            gathered[0] = comm.search_sharded(ix, xq_d.data_ptr(), n, K_TOP, stream.cuda_stream, nprobe=w["nprobe"],
                                              recall_num=RECALL_NUM, metric="L2", has_rank=True, deferred=deferred)
            return
        step_plain()
        if world > 1:
            dist.all_gather_into_tensor(out_all, out_d)

    # correctness side: recall@10 vs exact ground truth (and the raw result for the CPU cross-check)
    step_plain()
    torch.cuda.synchronize()
    I_ours = I_d.cpu().numpy().copy()
    valid_mask = None
    if filters:
        vm = filters[0][3].astype(bool).copy()
        vm[dele] = False
        valid_mask = torch.from_numpy(vm).to(dev)
    xb_t = torch.from_numpy(xb[: min(N, 10_000_000)]).to(dev) if (N <= 12_000_000 and not devb) else None
    rec_ours = None
    gt = None
    if xb_t is not None:
        gt = ground_truth(xb_t, xq_d, K_TOP, valid_mask).cpu().numpy()
        rec_ours = recall_at_k(I_ours, gt)
        del xb_t
        torch.cuda.empty_cache()
    elif devb:  # exact ground truth for the first 256 queries over regenerated rows
        nq_gt = min(256, n)
        gt = ground_truth_regenerated(devb["data"], N, xq_d[:nq_gt], K_TOP).cpu().numpy()
        rec_ours = recall_at_k(I_ours[:nq_gt], gt)
        torch.cuda.empty_cache()
    log("recall@10 (ours) = %s" % rec_ours)

    # ---- device-timed region: per-step CUDA events, L2 flushed between steps (untimed)
    ix.set_profiling(True)
    for _ in range(args.warmup):
        flush.fill_(1)
        step_dev()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local, gpu_uuid)
    sampler.start()

    def load_for(iters):  # the same step, untimed, a FIXED count on every rank (the step holds a collective when N > 1):
        for i_ in range(iters):  # the clock samples then see the load the timed region runs under
            flush.fill_(i_ & 0xff)
            step_dev()
        torch.cuda.synchronize()

    load_for(400)
    if world > 1:
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ix.launch_count()
    mark0 = sampler.mark()
    scan_ms, stage_acc = [], dict(coarse=0.0, scan=0.0, rerank=0.0, total=0.0)
    scanned = 0
    torch.cuda.synchronize()
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        evs[i][0].record(stream)
        step_dev()
        evs[i][1].record(stream)
        ix.sync()  # host sync so the per-stage event times of this step can be read
        st = ix.last_stage_ms()
        scan_ms.append(ix.last_scan_kernel_ms())
        for k_ in stage_acc:
            stage_acc[k_] += st[k_] / args.steps
        scanned = ix.last_scanned_postings()
    torch.cuda.synchronize()
    mark1 = sampler.mark()
    launches = ix.launch_count() - launches0
    if world > 1:
        dist.barrier()
    load_for(400)
    clocks = sampler.stop(mark0, mark1)
    if world > 1 and not p2p:
        launches += args.steps  # the NCCL all-gather per step (p2p: the re-rank kernel itself feeds the peers)
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    tm = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    ms_per_step = total_ms / args.steps
    value = world * n * args.steps / (total_ms / 1e3)

    # back-to-back, no flush (informational)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    torch.cuda.synchronize()
    ms_noflush = e0.elapsed_time(e1) / args.steps

    # ---- dev: tuning-knob sweep (stderr only, never part of the JSON line)
    for var in [v for v in args.variants.split(";") if v]:
        kv = dict(x.split("=") for x in var.split(","))
        for k_, v_ in kv.items():
            os.environ[k_] = v_
        ix.reload_tuning()
        for _ in range(3):
            step_dev()
        torch.cuda.synchronize()
        sc, sk = [], []
        t_ev = []
        for i in range(args.steps):
            flush.fill_(i & 0xff)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(stream)
            step_dev()
            b_.record(stream)
            ix.sync()
            sc.append(ix.last_stage_ms()["scan"])
            sk.append(ix.last_scan_kernel_ms())
            t_ev.append(a_.elapsed_time(b_))
        ok_ = bool(np.array_equal(I_d.cpu().numpy(), I_ours))
        log("variant %s: ms/step %.4f scan stage %.4f scan kernel %.4f same_ids=%s" % (
            var, float(np.mean(t_ev)), float(np.mean(sc)), float(np.mean(sk)), ok_))
        for k_ in kv:
            os.environ.pop(k_, None)
        ix.reload_tuning()

    # ---- e2e: public host call, pinned host buffers, H2D + D2H inside the timed region
    ix.set_profiling(False)  # no per-stage events on the path a user runs
    xq_pin = torch.from_numpy(xq).pin_memory()
    D_pin = torch.empty(n, K_TOP, dtype=torch.float32).pin_memory()
    I_pin = torch.empty(n, K_TOP, dtype=torch.int64).pin_memory()
    sp = api._Base._sp("L2", w["nprobe"], RECALL_NUM, True, -api.FLT_MAX, api.FLT_MAX)
    farr, fkeep = api.make_filters(filters)

    if world > 1:
        out_all_pin = torch.empty(world * out_bytes, dtype=torch.uint8).pin_memory()
        xq_stage = torch.empty_like(xq_d)

    def read_gathered(base):
        if base is None:
            stream.synchronize()
            return
        for r_ in range(world):  # rank r's block sits at r * slot_bytes in the window
            comm.read(out_all_pin.data_ptr() + r_ * out_bytes, base + r_ * comm.slot_bytes, out_bytes,
                      stream.cuda_stream, sync=(r_ == world - 1))

    def step_host():
        if world == 1:
            rc_ = api.lib().gb200_ivfpq_search(ix.h, n, xq_pin.data_ptr(), K_TOP, ctypes.byref(sp),
                                               ctypes.cast(farr, ctypes.c_void_p), len(filters), D_pin.data_ptr(),
                                               I_pin.data_ptr())
            assert rc_ == 0, api.lib().gb200_last_error()
            return
        # N > 1: the whole query-sharded path — H2D of this rank's queries, search, the all-gather of every rank's top-k
        # over NVLink, D2H of the gathered result — so that e2e contains the exchange
        xq_stage.copy_(xq_pin, non_blocking=True)
        if p2p:
            base = comm.search_sharded(ix, xq_stage.data_ptr(), n, K_TOP, stream.cuda_stream, nprobe=w["nprobe"],
                                       recall_num=RECALL_NUM, metric="L2", has_rank=True, deferred=deferred)
            read_gathered(base)  # deferred: the gathered result of the step before this one
            return
        rc_ = ix.search_dev(xq_stage.data_ptr(), n, K_TOP, D_d.data_ptr(), I_d.data_ptr(), stream.cuda_stream,
                            nprobe=w["nprobe"], recall_num=RECALL_NUM, metric="L2", has_rank=True)
        assert rc_ == 0, api.lib().gb200_last_error()
        dist.all_gather_into_tensor(out_all, out_d)
        out_all_pin.copy_(out_all, non_blocking=True)
        stream.synchronize()

    for _ in range(args.warmup):
        step_host()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    if deferred:  # the last step's gathered result: wait for it and read it inside the timed region
        read_gathered(comm.flush(stream.cuda_stream))
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_qps = world * n * args.steps / float(te.item())
    if world == 1:
        assert np.array_equal(I_pin.numpy(), I_ours), "host-API result differs from device-API result"
    else:
        mine = out_all_pin[rank * out_bytes:(rank + 1) * out_bytes][n * K_TOP * 4:].view(torch.int64).numpy().reshape(n, K_TOP)
        assert np.array_equal(mine, I_ours), "gathered result of this rank differs from its device-API result"

    if comm is not None:
        st_ = comm.status()
        assert st_ == 0, "exchange: peer %d did not arrive" % (st_ - 1)
        dist.barrier()
        comm.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (ADC scan)
    peak, peak_src = measured_peaks()
    code_size = w["M"]
    alg_bytes = scanned * (code_size + 4)
    scan_avg_ms = float(np.mean(scan_ms))
    achieved = alg_bytes / (scan_avg_ms / 1e3) / 1e9
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=None,
                    kernel=("ivfpq_scan_m32_v3_kernel" if w["M"] == 32 else
                            "ivfpq_scan_m64_kernel" if w["M"] == 64 else
                            "ivfpq_scan_generic_kernel") + " (CUDA events around that one launch on the search stream)",
                    algorithmic_bytes_per_launch=alg_bytes,
                    scanned_postings_per_launch=scanned, kernel_ms=scan_avg_ms, peak_source=peak_src,
                    stage_ms=stage_acc)
    tr = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(args.workload)
        except Exception:
            pass

    # ---- CPU baseline: reference engine, same index, bounded sample
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            r = build_reference(w, N_cpu, nlist, xb, coarse, pq, list_no, codes, dele)
            nq = min(args.cpu_queries, n)
            qps, cores, times, I_cpu = time_reference(r, w, xq, filters, nq, reps=3)
            cpu = dict(value=qps, unit="queries/s", cores=cores, kind="reference",
                       sample="first %d queries of the batch, median of 3 runs after 1 warm-up, parallel_on_queries=1%s" % (nq, cpu_note))
            if not devb:  # same index on both sides: the ids must be the same
                cpu["ids_identical_frac"] = float((I_cpu == I_ours[:nq]).mean())
                if rec_ours is not None:
                    cpu["recall_at_10"] = recall_at_k(I_cpu, gt[:nq])
            r.close()
        except Exception as e:  # the baseline is reported, never required for the GPU number
            cpu = dict(value=None, unit="queries/s", cores=None, kind="reference", sample="failed: %r" % (e,))

    out = dict(metric=w.get("metric", METRIC), value=value, unit="queries/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
               ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
               data="synthetic",
               config=dict(workload=w["desc"], N=N, nlist=nlist, M=w["M"], nprobe=w["nprobe"], batch_per_gpu=n,
                           global_batch=world * n, k=K_TOP, recall_num=RECALL_NUM, has_rank=True,
                           parallelism="query-sharded x%d, index replicated, %s" % (
                               world, ("top-k pushed into every peer's window over NVLink, a step waits for the peers' results "
                                       "of the step before it (gb200_ivfpq_search_sharded_deferred)") if deferred else
                               "top-k pushed into every peer's window over NVLink (gb200_ivfpq_search_sharded)" if p2p
                               else "NCCL all-gather of top-k"),
                           l2="256 MB flush write between steps (untimed); per-step CUDA events",
                           ms_per_step_back_to_back_no_flush=ms_noflush, scaled_down=args.scale != 1.0),
               roofline=roofline, cpu_baseline=cpu,
               e2e=dict(value=e2e_qps, unit="queries/s", h2d_bytes_per_step=int(n * w["d"] * 4),
                        d2h_bytes_per_step=int(n * K_TOP * 12 * (world if world > 1 else 1)),
                        path=("gb200_ivfpq_search (host C-ABI, pinned host buffers)" if world == 1 else
                              "per rank: H2D queries, search + exchange (%s), D2H of the gathered result" % (
                                  "gb200_ivfpq_search_sharded_deferred, result read one step late, last one after gb200_comm_flush"
                                  if deferred else "gb200_ivfpq_search_sharded" if p2p else
                                  "gb200_ivfpq_search_dev + NCCL all-gather"))),
               gpu_launches=int(launches), clocks=clocks, recall_at_10=rec_ours)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
