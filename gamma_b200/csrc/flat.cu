// K4a — exact flat scan over the raw-vector mirror (CUDA-core path; bit-exact distances).
// Replaces GammaFLATIndex::Search / search_impl (index/impl/gamma_index_flat.cc:118-300):
//   for vid in [0, N): skip !IsValid(vid); dis = fvec_L2sqr / fvec_inner_product(xi, raw[vid], d);
//   skip if outside [min_score, max_score]; keep the k best (strict compare => first scanned wins);
//   heap_reorder.
// Work split: grid = (splits of the vid range, queries); every CTA keeps its k best in a
// BlockTopR, a second tiny kernel merges the splits.  This path streams the database once PER
// QUERY, so it is the small-batch / parity path; large batches go through the tensor-core
// kernel in flat_tc.cu, which uses this file's exact_distance for its final re-score.
#include "common.cuh"
#include "kernels.h"

namespace gb {

constexpr int FL_THREADS = 256;
constexpr int FL_OCT = FL_THREADS / 8;  // candidates per pass
constexpr int FL_U = 4;                 // passes per round

template <bool IP, int PER>
__global__ void __launch_bounds__(FL_THREADS) flat_exact_kernel(FlatParams P, int cap, int kpad) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64 *buf = reinterpret_cast<u64 *>(smem);
  int *misc = reinterpret_cast<int *>(smem + (size_t)cap * sizeof(u64));
  float *qs = reinterpret_cast<float *>(misc + 4 + 64);
  BlockTopR topr;
  topr.buf = buf;
  topr.tau = reinterpret_cast<u64 *>(misc);
  topr.cnt = misc + 2;
  topr.warp_part = misc + 4;
  topr.cap = cap;
  topr.R = P.k;
  const int q = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < P.d; i += FL_THREADS) qs[i] = P.xq[(size_t)q * P.d + i];
  topr.init_collective();

  const long long per = (P.N + P.nsplit - 1) / P.nsplit;
  const long long v0 = (long long)split * per;
  const long long v1 = min(P.N, v0 + per);
  const int sub = tid & 7, oct = tid >> 3;
  const int prune_limit = cap - FL_OCT * FL_U;
  for (long long base = v0; base < v1; base += FL_OCT * FL_U) {
#pragma unroll
    for (int u = 0; u < FL_U; u++) {
      long long vid = base + u * FL_OCT + oct;
      bool ok = vid < v1;
      if (ok && P.valid) ok = bitmap_test(P.valid, (int)vid);
      const float *y = P.raw + (size_t)(ok ? vid : v0) * P.d;
      float dis = exact_distance_octet<IP>(qs, y, ok ? P.d : 0, sub);
      ok = ok && dis <= P.max_score && dis >= P.min_score;
      u64 key = ((u64)dist_to_key32<IP>(dis) << 32) | (uint32_t)vid;
      bool pass = ok && sub == 0 && key < topr.threshold();
      topr.append_warp(pass, key);
    }
    int over = *((volatile int *)topr.cnt) > prune_limit;
    if (__syncthreads_or(over)) topr.prune_collective<PER>();
  }
  topr.prune_collective<PER>();
  const int n_out = min(*((volatile int *)topr.cnt), P.k);
  u64 *out = P.scratch + ((size_t)q * P.nsplit + split) * kpad;
  for (int i = tid; i < kpad; i += FL_THREADS) out[i] = i < n_out ? buf[i] : GB_KEY_MAX;
}

template <bool IP>
__global__ void __launch_bounds__(256) flat_merge_kernel(FlatParams P, int kpad, int p2) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64 *keys = reinterpret_cast<u64 *>(smem);
  const int q = blockIdx.x, tid = threadIdx.x;
  const int total = P.nsplit * kpad;
  const u64 *src = P.scratch + (size_t)q * total;
  for (int i = tid; i < p2; i += 256) keys[i] = i < total ? src[i] : GB_KEY_MAX;
  __syncthreads();
  block_bitonic_sort(keys, p2);
  const float neutral = IP ? -3.402823466e38f : 3.402823466e38f;
  for (int j = tid; j < P.k; j += 256) {
    u64 k = keys[j];
    bool have = k != GB_KEY_MAX;
    P.out_dist[(size_t)q * P.k + j] = have ? key32_to_dist<IP>((uint32_t)(k >> 32)) : neutral;
    P.out_ids[(size_t)q * P.k + j] = have ? (long long)(uint32_t)k : -1;
  }
}

int flat_exact_splits(long long N, int n) {
  // enough CTAs to fill 148 SMs a few times over, but keep nsplit * kpad mergeable
  long long want = (148 * 4 + n - 1) / n;
  long long by_work = (N + 4095) / 4096;
  long long s = want < by_work ? want : by_work;
  if (s < 1) s = 1;
  if (s > 512) s = 512;
  return (int)s;
}

cudaError_t launch_flat_exact(const FlatParams &P, cudaStream_t st) {
  int kpad = P.k;
  int need = P.k + FL_OCT * FL_U;
  int cap = 512;
  while (cap < need) cap <<= 1;
  size_t smem = (size_t)cap * sizeof(u64) + (4 + 64) * sizeof(int) + (size_t)P.d * sizeof(float);
  int p2 = next_pow2(P.nsplit * kpad);
  if (p2 < next_pow2(P.k)) p2 = next_pow2(P.k);
  size_t smem2 = (size_t)p2 * sizeof(u64);
  if (smem > 200 * 1024 || smem2 > 200 * 1024) return cudaErrorInvalidValue;
  auto run = [&](auto exact, auto merge) -> cudaError_t {
    cudaError_t e;
    e = ensure_dynamic_smem(exact, smem);
    if (e) return e;
    e = ensure_dynamic_smem(merge, smem2);
    if (e) return e;
    exact<<<dim3(P.nsplit, P.n), FL_THREADS, smem, st>>>(P, cap, kpad);
    merge<<<P.n, 256, smem2, st>>>(P, kpad, p2);
    return cudaGetLastError();
  };
  const bool big = cap > 4 * FL_THREADS;
  if (P.is_ip) return big ? run(flat_exact_kernel<true, 16>, flat_merge_kernel<true>) : run(flat_exact_kernel<true, 4>, flat_merge_kernel<true>);
  return big ? run(flat_exact_kernel<false, 16>, flat_merge_kernel<false>) : run(flat_exact_kernel<false, 4>, flat_merge_kernel<false>);
}

}  // namespace gb
