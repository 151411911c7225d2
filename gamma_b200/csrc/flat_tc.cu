// K4t — batched FLAT search through the tensor cores.
// GammaFLATIndex::Search (index/impl/gamma_index_flat.cc:118-300) evaluates n x N exact distances; for a
// batch this is a dense Q x Y^T contraction (BASELINE config 4: IP, d = 768, 5M vectors, batch 512 =
// 3.9 TFLOP), so the distance PRODUCER is the tcgen05 3xTF32 GEMM of tc_gemm.cu, chunk by chunk over the
// database.  Everything that defines the result stays exact:
//   1. per chunk: distances of the chunk (fp32-level accuracy) -> this file's running select keeps the
//      K' = k + margin best (key = distance, vid) per query with the validity bitmap and a slightly
//      widened score window applied — candidates only;
//   2. after the last chunk: every candidate is re-scored with exact_distance (the AVX-order fp32
//      kernel of flat.cu/rerank.cu), the exact score window is applied, and the k best are emitted in
//      (distance, vid) order — the same values and order the CPU engine produces.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace gb {

constexpr int FT_THREADS = 256;
constexpr int FT_PER_ROUND = 4;

// one CTA per query: stream the query's row of the chunk's distance tile, keep the Kp best seen so far
template <bool IP, int PER>
__global__ void __launch_bounds__(FT_THREADS) flat_chunk_select_kernel(const float *__restrict__ dist, int ldo, int nc,
                                                                       long long chunk_base,
                                                                       const uint32_t *__restrict__ valid, float lo,
                                                                       float hi, int Kp, int cap, int first,
                                                                       u64 *__restrict__ state) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64 *buf = reinterpret_cast<u64 *>(smem);
  int *misc = reinterpret_cast<int *>(smem + (size_t)cap * sizeof(u64));
  BlockTopR topr;
  topr.buf = buf;
  topr.tau = reinterpret_cast<u64 *>(misc);
  topr.cnt = misc + 2;
  topr.warp_part = misc + 4;
  topr.cap = cap;
  topr.R = Kp;
  const int q = blockIdx.x, tid = threadIdx.x;
  u64 *st = state + (size_t)q * Kp;
  // seed with the survivors of the previous chunks
  int seeded = 0;
  if (!first) {
    for (int i = tid; i < Kp; i += FT_THREADS) buf[i] = st[i];
    seeded = Kp;
  }
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    if (!first)
      while (c < Kp && buf[c] != GB_KEY_MAX) c++;  // survivors are packed at the front
    *topr.cnt = c;
    *topr.tau = GB_KEY_MAX;
  }
  (void)seeded;
  __syncthreads();
  const float *row = dist + (size_t)q * ldo;
  const int per_round = FT_THREADS * FT_PER_ROUND;
  const int prune_limit = cap - per_round;
  if (*((volatile int *)topr.cnt) > prune_limit) topr.prune_collective<PER>();  // uniform
  for (int base = 0; base < nc; base += per_round) {
#pragma unroll
    for (int t = 0; t < FT_PER_ROUND; t++) {
      int j = base + t * FT_THREADS + tid;
      bool ok = j < nc;
      float v = ok ? row[j] : 0.f;
      long long vid = chunk_base + j;
      ok = ok && (v == v) && v >= lo && v <= hi;
      if (ok && valid) ok = bitmap_test(valid, (int)vid);
      u64 key = ((u64)dist_to_key32<IP>(v) << 32) | (uint32_t)vid;
      bool pass = ok && key < topr.threshold();
      topr.append_warp(pass, key);
    }
    int over = *((volatile int *)topr.cnt) > prune_limit;
    if (__syncthreads_or(over)) topr.prune_collective<PER>();
  }
  topr.prune_collective<PER>();
  const int n_out = min(*((volatile int *)topr.cnt), Kp);
  for (int i = tid; i < Kp; i += FT_THREADS) st[i] = i < n_out ? buf[i] : GB_KEY_MAX;
}

// Streamlined variant (the one launched): after the first chunk the running threshold rejects almost every
// element, so the common path is one 16-byte load per 4 elements, a key transform and a 32-bit compare; only the
// rare survivor pays for the 64-bit test, the bitmap probe and a shared atomic.  The next round's loads are in
// flight while the current round is filtered, and the CTA meets once per 2048 elements (the round) to decide
// about a prune.  r01b: the per-element ballot + barrier version above took 598 us per 512 x 131072 tile
// against 432 us for the GEMM that produced it.
constexpr int FT2_CAP = 4096;            // 16 keys per thread in the select
constexpr int FT2_ROUND = FT_THREADS * 8;  // elements per round: 2 x float4 per thread

template <bool IP>
__global__ void __launch_bounds__(FT_THREADS) flat_chunk_select2_kernel(const float *__restrict__ dist, int ldo, int nc,
                                                                        long long chunk_base,
                                                                        const uint32_t *__restrict__ valid, float lo,
                                                                        float hi, int Kp, int first,
                                                                        u64 *__restrict__ state) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64 *buf = reinterpret_cast<u64 *>(smem);
  int *misc = reinterpret_cast<int *>(smem + (size_t)FT2_CAP * sizeof(u64));
  BlockTopR topr;
  topr.buf = buf;
  topr.tau = reinterpret_cast<u64 *>(misc);
  topr.cnt = misc + 2;
  topr.warp_part = misc + 4;
  topr.cap = FT2_CAP;
  topr.R = Kp;
  const int q = blockIdx.x, tid = threadIdx.x;
  u64 *st = state + (size_t)q * Kp;
  // seed with the survivors of the previous chunks (packed at the front); a full state gives the threshold
  u64 mx = 0;
  int have = 0;
  if (!first)
    for (int i = tid; i < Kp; i += FT_THREADS) {
      const u64 k = st[i];
      buf[i] = k;
      if (k != GB_KEY_MAX) {
        have++;
        mx = k > mx ? k : mx;
      }
    }
  have = __reduce_add_sync(GB_FULL, have);
  for (int o = 16; o; o >>= 1) {
    const u64 x = __shfl_xor_sync(GB_FULL, mx, o);
    mx = x > mx ? x : mx;
  }
  u64 *wmax = reinterpret_cast<u64 *>(misc + 20);  // [8] scratch inside warp_part
  if ((tid & 31) == 0) {
    misc[4 + (tid >> 5)] = have;
    wmax[tid >> 5] = mx;
  }
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    u64 m = 0;
    for (int w = 0; w < FT_THREADS / 32; w++) {
      c += misc[4 + w];
      m = wmax[w] > m ? wmax[w] : m;
    }
    *topr.cnt = c;
    *topr.tau = c >= Kp ? m + 1 : GB_KEY_MAX;  // full: only keys better than the current worst can enter
  }
  __syncthreads();
  const float *row = dist + (size_t)q * ldo;
  const int prune_limit = FT2_CAP - FT2_ROUND;
  const float4 none = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load = [&](int base, int t) -> float4 {
    const int j = base + (t * FT_THREADS + tid) * 4;
    return j < nc ? __ldg(reinterpret_cast<const float4 *>(row + j)) : none;  // nc % 4 == 0 (host checks)
  };
  float4 a = load(0, 0), b = load(0, 1);
  for (int base = 0; base < nc; base += FT2_ROUND) {
    const float4 ca = a, cb = b;
    a = load(base + FT2_ROUND, 0), b = load(base + FT2_ROUND, 1);  // next round in flight
    const u64 tau = topr.threshold();
    const uint32_t tau_hi = (uint32_t)(tau >> 32);
    auto visit = [&](float v, int j) {
      const uint32_t k32 = dist_to_key32<IP>(v);
      if (j < nc && k32 <= tau_hi && v >= lo && v <= hi) {  // NaN fails the window test
        const long long vid = chunk_base + j;
        const u64 key = ((u64)k32 << 32) | (uint32_t)vid;
        if (key < tau && (!valid || bitmap_test(valid, (int)vid))) {
          const int slot = atomicAdd(topr.cnt, 1);
          if (slot < FT2_CAP) buf[slot] = key;  // cannot overflow: cnt <= prune_limit at the start of a round
        }
      }
    };
    const int j0 = base + tid * 4, j1 = base + (FT_THREADS + tid) * 4;
    visit(ca.x, j0), visit(ca.y, j0 + 1), visit(ca.z, j0 + 2), visit(ca.w, j0 + 3);
    visit(cb.x, j1), visit(cb.y, j1 + 1), visit(cb.z, j1 + 2), visit(cb.w, j1 + 3);
    const int over = *((volatile int *)topr.cnt) > prune_limit;
    if (__syncthreads_or(over)) topr.prune_collective<16>();
  }
  topr.prune_collective<16>();
  const int n_out = min(*((volatile int *)topr.cnt), Kp);
  for (int i = tid; i < Kp; i += FT_THREADS) st[i] = i < n_out ? buf[i] : GB_KEY_MAX;
}


// exact re-score of the Kp candidates of every query, exact score window, k best in (distance, vid) order
template <bool IP>
__global__ void __launch_bounds__(128) flat_rescore_kernel(const u64 *__restrict__ state, int Kp, int p2,
                                                           const float *__restrict__ xq, const float *__restrict__ raw,
                                                           int d, float min_score, float max_score, int k,
                                                           float *__restrict__ out_d, long long *__restrict__ out_i) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64 *keys = reinterpret_cast<u64 *>(smem);           // [p2]
  float *qs = reinterpret_cast<float *>(keys + p2);    // [d]
  const int q = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < d; i += 128) qs[i] = xq[(size_t)q * d + i];
  for (int i = tid; i < p2; i += 128) keys[i] = GB_KEY_MAX;
  __syncthreads();
  const int sub = tid & 7, oct = tid >> 3;
  for (int base = 0; base < Kp; base += 16) {
    int i = base + oct;
    u64 ck = i < Kp ? state[(size_t)q * Kp + i] : GB_KEY_MAX;
    bool have = ck != GB_KEY_MAX;
    uint32_t vid = (uint32_t)ck;
    const float *y = raw + (size_t)(have ? vid : 0) * d;
    float dis = exact_distance_octet<IP>(qs, y, have ? d : 0, sub);
    if (sub == 0 && i < Kp) {
      bool ok = have && dis <= max_score && dis >= min_score;  // IsSimilarScoreValid on the exact value
      keys[i] = ok ? (((u64)dist_to_key32<IP>(dis) << 32) | vid) : GB_KEY_MAX;
    }
  }
  __syncthreads();
  block_bitonic_sort(keys, p2);
  const float neutral = IP ? -3.402823466e38f : 3.402823466e38f;
  for (int j = tid; j < k; j += 128) {
    u64 kk = j < p2 ? keys[j] : GB_KEY_MAX;
    bool have = kk != GB_KEY_MAX;
    out_d[(size_t)q * k + j] = have ? key32_to_dist<IP>((uint32_t)(kk >> 32)) : neutral;
    out_i[(size_t)q * k + j] = have ? (long long)(uint32_t)kk : -1;
  }
}

int flat_tc_candidates(int k) { return k + 64; }

cudaError_t launch_flat_chunk_select(const float *dist, int ldo, int nc, long long chunk_base, const uint32_t *valid,
                                     float lo, float hi, int Kp, int first, u64 *state, int n, int is_ip,
                                     cudaStream_t st) {
  if (Kp <= FT2_CAP - FT2_ROUND && (ldo & 3) == 0 && (nc & 3) == 0 && ((uintptr_t)dist & 15) == 0) {
    const size_t smem2 = (size_t)FT2_CAP * sizeof(u64) + (4 + 64) * sizeof(int);
    if (is_ip)
      flat_chunk_select2_kernel<true><<<n, FT_THREADS, smem2, st>>>(dist, ldo, nc, chunk_base, valid, lo, hi, Kp, first, state);
    else
      flat_chunk_select2_kernel<false><<<n, FT_THREADS, smem2, st>>>(dist, ldo, nc, chunk_base, valid, lo, hi, Kp, first, state);
    return cudaGetLastError();
  }
  int need = Kp + FT_THREADS * FT_PER_ROUND;
  int cap = 1024;
  while (cap < need) cap <<= 1;
  size_t smem = (size_t)cap * sizeof(u64) + (4 + 64) * sizeof(int);
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e) return e;
    kern<<<n, FT_THREADS, smem, st>>>(dist, ldo, nc, chunk_base, valid, lo, hi, Kp, cap, first, state);
    return cudaGetLastError();
  };
  const bool big = cap > 4 * FT_THREADS;
  if (is_ip) return big ? go(flat_chunk_select_kernel<true, 16>) : go(flat_chunk_select_kernel<true, 4>);
  return big ? go(flat_chunk_select_kernel<false, 16>) : go(flat_chunk_select_kernel<false, 4>);
}

cudaError_t launch_flat_rescore(const u64 *state, int Kp, const float *xq, const float *raw, int n, int d,
                                float min_score, float max_score, int k, int is_ip, float *out_d, long long *out_i,
                                cudaStream_t st) {
  int p2 = next_pow2(Kp > k ? Kp : k);
  size_t smem = (size_t)p2 * sizeof(u64) + (size_t)d * sizeof(float);
  if (is_ip) {
    cudaError_t e = ensure_dynamic_smem(flat_rescore_kernel<true>, smem);
    if (e) return e;
    flat_rescore_kernel<true><<<n, 128, smem, st>>>(state, Kp, p2, xq, raw, d, min_score, max_score, k, out_d, out_i);
  } else {
    cudaError_t e = ensure_dynamic_smem(flat_rescore_kernel<false>, smem);
    if (e) return e;
    flat_rescore_kernel<false><<<n, 128, smem, st>>>(state, Kp, p2, xq, raw, d, min_score, max_score, k, out_d, out_i);
  }
  return cudaGetLastError();
}

}  // namespace gb
