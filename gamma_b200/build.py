"""Build libgamma_b200.so (hand-written sm_100a CUDA + the extern "C" shim) in-tree with nvcc.

    python -m gamma_b200.build        # -> gamma_b200/lib/libgamma_b200.so

nvcc cross-compiles for sm_100a without a GPU; the built .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libgamma_b200.so")
SOURCES = ["capi.cu", "ivfpq_scan.cu", "ivfpq_scan_v3.cu", "postings.cu", "coarse.cu", "rerank.cu", "flat.cu", "selftest.cu", "tc_gemm.cu", "flat_tc.cu", "encode.cu", "comm.cu", "ivfflat.cu", "launch.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "gamma_b200.h"))
    objs, procs = [], []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(op)
        if force or _newer(sp, op) or any(_newer(h, op) for h in headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", op]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write("== nvcc %s\n%s\n" % (src, out))
        with open(os.path.join(objdir, src + ".ptxas.log"), "w") as f:
            f.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
