// Test hook: run BlockTopR (append + radix-select prune) on caller-provided keys and return the
// survivors, so tests can check the selection primitive against a host sort in isolation.
#include "common.cuh"
#include "kernels.h"

namespace gb {

template <int PER>
__global__ void select_selftest_kernel(const u64 *keys, int n, int R, int cap, int batch, u64 *out, int *out_n) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64 *buf = reinterpret_cast<u64 *>(smem);
  int *misc = reinterpret_cast<int *>(smem + (size_t)cap * sizeof(u64));
  BlockTopR topr;
  topr.buf = buf;
  topr.tau = reinterpret_cast<u64 *>(misc);
  topr.cnt = misc + 2;
  topr.warp_part = misc + 4;
  topr.cap = cap;
  topr.R = R;
  topr.init_collective();
  // feed in batches of `batch` keys (<= cap - R), pruning whenever the next batch might not fit
  for (int base = 0; base < n; base += batch) {
    if (*((volatile int *)topr.cnt) + batch > cap) topr.prune_collective<PER>();  // uniform: cnt read after a barrier
    __syncthreads();
    for (int i = base + threadIdx.x; i < base + batch; i += blockDim.x) {
      // warp-collective append: all lanes of a warp must call it
    }
    int iters = (batch + blockDim.x - 1) / blockDim.x;
    for (int it = 0; it < iters; it++) {
      int i = base + it * blockDim.x + threadIdx.x;
      bool ok = i < n && i < base + batch;
      u64 key = ok ? keys[i] : GB_KEY_MAX;
      topr.append_warp(ok && key < topr.threshold(), key);
    }
    __syncthreads();
  }
  topr.prune_collective<PER>();
  int m = min(*((volatile int *)topr.cnt), R);
  for (int i = threadIdx.x; i < m; i += blockDim.x) out[i] = buf[i];
  if (threadIdx.x == 0) *out_n = m;
}

cudaError_t launch_select_selftest(const u64 *keys, int n, int R, int cap, int batch, int threads, u64 *out, int *out_n,
                                   cudaStream_t st) {
  size_t smem = (size_t)cap * sizeof(u64) + (4 + 64) * sizeof(int);
  cudaError_t e;
  if (cap <= 4 * threads) {
    e = ensure_dynamic_smem(select_selftest_kernel<4>, smem);
    if (e) return e;
    select_selftest_kernel<4><<<1, threads, smem, st>>>(keys, n, R, cap, batch, out, out_n);
  } else {
    e = ensure_dynamic_smem(select_selftest_kernel<16>, smem);
    if (e) return e;
    select_selftest_kernel<16><<<1, threads, smem, st>>>(keys, n, R, cap, batch, out, out_n);
  }
  return cudaGetLastError();
}

}  // namespace gb
