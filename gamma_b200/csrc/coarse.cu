// K1 — coarse quantiser: the nprobe nearest coarse centroids of every query, ascending L2^2.
// Replaces quantizer->search (index/impl/gamma_index_ivfpq.cc:560) = faiss::IndexFlatL2::search
// -> knn_L2sqr, BLAS path (faiss utils/distances.cpp:215-296):
//     dis(i,j) = |x_i|^2 + |y_j|^2 - 2 <x_i, y_j>,  negative values clamped to 0,
// followed by a k-select of the nprobe smallest.  The inner product is accumulated in fp32
// (FMA), so the probe set matches the CPU engine up to rounding-level ties in coarse distance.
#include "common.cuh"
#include "kernels.h"

namespace gb {

constexpr int CT = 64;   // tile (queries x centroids)
constexpr int CK = 16;   // k-slab

__global__ void __launch_bounds__(256) coarse_dist_kernel(const float *__restrict__ xq, const float *__restrict__ xn,
                                                          const float *__restrict__ cent,
                                                          const float *__restrict__ cn, int n, int nlist, int d,
                                                          float *__restrict__ dist) {
  __shared__ float As[CK][CT + 4];
  __shared__ float Bs[CK][CT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4x4 outputs each
  const int row0 = blockIdx.y * CT, col0 = blockIdx.x * CT;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < d; k0 += CK) {
    // 64 rows x 16 k per operand = 1024 floats, 4 per thread
#pragma unroll
    for (int t = 0; t < 4; t++) {
      int e = threadIdx.x + t * 256;
      int r = e >> 4, kk = e & 15;
      int gk = k0 + kk;
      int ga = row0 + r, gb_ = col0 + r;
      As[kk][r] = (ga < n && gk < d) ? xq[(size_t)ga * d + gk] : 0.f;
      Bs[kk][r] = (gb_ < nlist && gk < d) ? cent[(size_t)gb_ * d + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CK; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int r = row0 + ty * 4 + i;
    if (r >= n) continue;
    float qn = xn[r];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int c = col0 + tx * 4 + j;
      if (c >= nlist) continue;
      float v = qn + cn[c] - 2.f * acc[i][j];
      dist[(size_t)r * nlist + c] = v < 0.f ? 0.f : v;
    }
  }
}

cudaError_t launch_coarse_dist(const float *xq, const float *xq_norm, const float *cent, const float *cent_norm,
                               int n, int nlist, int d, float *dist, cudaStream_t st) {
  dim3 grid((nlist + CT - 1) / CT, (n + CT - 1) / CT);
  coarse_dist_kernel<<<grid, 256, 0, st>>>(xq, xq_norm, cent, cent_norm, n, nlist, d, dist);
  return cudaGetLastError();
}

constexpr int CS_THREADS = 256;
constexpr int CS_PER_ROUND = 4;  // elements per thread per round

template <int PER>
__global__ void __launch_bounds__(CS_THREADS) coarse_select_kernel(const float *__restrict__ dist, int nlist,
                                                                    int nprobe, int cap, int *__restrict__ keys,
                                                                    float *__restrict__ coarse_dis) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64 *buf = reinterpret_cast<u64 *>(smem);
  int *misc = reinterpret_cast<int *>(smem + (size_t)cap * sizeof(u64));
  BlockTopR topr;
  topr.buf = buf;
  topr.tau = reinterpret_cast<u64 *>(misc);
  topr.cnt = misc + 2;
  topr.warp_part = misc + 4;
  topr.cap = cap;
  topr.R = nprobe;
  topr.init_collective();

  const int q = blockIdx.x;
  const float *row = dist + (size_t)q * nlist;
  const int per_round = CS_THREADS * CS_PER_ROUND;
  const int prune_limit = cap - per_round;
  for (int base = 0; base < nlist; base += per_round) {
#pragma unroll
    for (int t = 0; t < CS_PER_ROUND; t++) {
      int j = base + t * CS_THREADS + threadIdx.x;
      bool ok = j < nlist;
      float v = ok ? row[j] : 0.f;
      u64 key = ((u64)float_to_ordered(v) << 32) | (uint32_t)j;
      bool pass = ok && (v == v) && key < topr.threshold();
      topr.append_warp(pass, key);
    }
    int over = *((volatile int *)topr.cnt) > prune_limit;
    if (__syncthreads_or(over)) topr.prune_collective<PER>();
  }
  topr.prune_collective<PER>();
  const int n_out = min(*((volatile int *)topr.cnt), nprobe);
  const int np2 = next_pow2(nprobe);
  for (int i = n_out + threadIdx.x; i < np2; i += CS_THREADS) buf[i] = GB_KEY_MAX;
  __syncthreads();
  block_bitonic_sort(buf, np2);
  for (int i = threadIdx.x; i < nprobe; i += CS_THREADS) {
    u64 k = buf[i];
    bool have = i < n_out;
    keys[(size_t)q * nprobe + i] = have ? (int)(uint32_t)k : -1;  // "not enough centroids": key -1
    coarse_dis[(size_t)q * nprobe + i] = have ? ordered_to_float((uint32_t)(k >> 32)) : 3.402823466e38f;
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-per-row two-pass select for nprobe <= 128 (the common case): no sorting network, 4 barriers.
//   pass 1: every thread takes the minimum key of its column stripe; threads are folded into
//           G >= nprobe groups (G = 32/64/128), tau0 = the largest of the G group minima.  G distinct
//           keys are <= tau0, so the nprobe-th smallest key of the row is <= tau0
//           (keys = (ordered distance, column) are unique).
//   pass 2: keys <= tau0 (about G ln G of them in expectation) are appended to a shared buffer.
//   rank  : each survivor counts how many survivors are smaller; rank < nprobe writes output slot
//           `rank` — ascending distance, ties by centroid id, for free.
// If the survivors overflow the buffer (adversarial column order) tau is narrowed by bisection on
// the key space, re-counting over the row (slow, rare, exact).
// ---------------------------------------------------------------------------------------------
constexpr int CW_THREADS = 256;
constexpr int CW_CAP = 2048;  // survivors

__device__ __forceinline__ u64 cw_key(float v, int col) {
  return (v == v) ? (((u64)float_to_ordered(v) << 32) | (uint32_t)col) : GB_KEY_MAX;
}

__global__ void __launch_bounds__(CW_THREADS) coarse_select_row_kernel(const float *__restrict__ dist, int nlist, int nprobe,
                                                                       int G, int *__restrict__ keys,
                                                                       float *__restrict__ coarse_dis) {
  __shared__ u64 cand[CW_CAP];
  __shared__ u64 gmin[CW_THREADS];
  __shared__ u64 s_tau;
  __shared__ int s_cnt;
  const int tid = threadIdx.x, lane = tid & 31;
  const int q = blockIdx.x;
  const float *row = dist + (size_t)q * nlist;
  const bool vec4 = (nlist & 3) == 0 && ((uintptr_t)row & 15) == 0;

  // ---- pass 1: per-thread minimum
  u64 best = GB_KEY_MAX;
  if (vec4) {
    for (int c = tid * 4; c < nlist; c += CW_THREADS * 4) {
      float4 v = *reinterpret_cast<const float4 *>(row + c);
      u64 k0 = cw_key(v.x, c), k1 = cw_key(v.y, c + 1), k2 = cw_key(v.z, c + 2), k3 = cw_key(v.w, c + 3);
      k0 = k0 < k1 ? k0 : k1;
      k2 = k2 < k3 ? k2 : k3;
      k0 = k0 < k2 ? k0 : k2;
      best = best < k0 ? best : k0;
    }
  } else {
    for (int c = tid; c < nlist; c += CW_THREADS) {
      u64 k = cw_key(row[c], c);
      best = best < k ? best : k;
    }
  }
  gmin[tid] = best;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  if (tid < 32) {  // group g = tid % G : fold the 256 thread minima into G group minima, tau0 = their maximum
    u64 t = 0;
    for (int g = tid; g < G; g += 32) {
      u64 m = GB_KEY_MAX;
      for (int j = g; j < CW_THREADS; j += G) m = gmin[j] < m ? gmin[j] : m;
      t = m > t ? m : t;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      u64 x = __shfl_xor_sync(GB_FULL, t, o);
      t = x > t ? x : t;
    }
    if (tid == 0) s_tau = t;
  }
  __syncthreads();
  u64 tau = s_tau;
  const bool all_valid_pass = tau == GB_KEY_MAX;  // some group saw no usable key: everything valid survives

  // ---- pass 2 (+ bisection when the buffer overflows)
  u64 lo_b = 0, hi_b = tau;  // count(key <= hi_b) >= nprobe, count(key < lo_b) < nprobe
  int cnt = 0;
  for (int attempt = 0;; attempt++) {
    auto visit = [&](u64 key) {
      bool pass = key <= tau && key != GB_KEY_MAX;
      unsigned m = __ballot_sync(GB_FULL, pass);
      if (m) {
        int base = 0;
        int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(&s_cnt, __popc(m));
        base = __shfl_sync(GB_FULL, base, leader);
        int slot = base + __popc(m & ((1u << lane) - 1u));
        if (pass && slot < CW_CAP) cand[slot] = key;
      }
    };
    if (vec4) {
      const int lim = ((nlist + CW_THREADS * 4 - 1) / (CW_THREADS * 4)) * (CW_THREADS * 4);
      for (int c = tid * 4; c < lim; c += CW_THREADS * 4) {
        const bool in = c < nlist;
        float4 v = in ? *reinterpret_cast<const float4 *>(row + c) : make_float4(0, 0, 0, 0);
        visit(in ? cw_key(v.x, c) : GB_KEY_MAX);
        visit(in ? cw_key(v.y, c + 1) : GB_KEY_MAX);
        visit(in ? cw_key(v.z, c + 2) : GB_KEY_MAX);
        visit(in ? cw_key(v.w, c + 3) : GB_KEY_MAX);
      }
    } else {
      const int lim = ((nlist + CW_THREADS - 1) / CW_THREADS) * CW_THREADS;
      for (int c = tid; c < lim; c += CW_THREADS) visit(c < nlist ? cw_key(row[c], c) : GB_KEY_MAX);
    }
    __syncthreads();
    cnt = s_cnt;
    if (cnt <= CW_CAP && (cnt >= nprobe || (all_valid_pass && attempt == 0))) break;
    if (attempt > 130) break;  // cannot happen (the interval halves every step); guards against a hang
    if (cnt > CW_CAP)
      hi_b = tau;      // still >= nprobe survivors at tau: a valid upper bound, try lower
    else
      lo_b = tau + 1;  // fewer than nprobe survivors: tau was lowered too far
    tau = lo_b + (hi_b - lo_b) / 2;
    __syncthreads();
    if (tid == 0) s_cnt = 0;
    __syncthreads();
  }
  const int c = min(cnt, CW_CAP);
  // ---- rank by counting; survivors are unique keys
  for (int i = tid; i < c; i += CW_THREADS) {
    const u64 k = cand[i];
    int r = 0;
    for (int j = 0; j < c; j++) r += cand[j] < k;
    if (r < nprobe) {
      keys[(size_t)q * nprobe + r] = (int)(uint32_t)k;
      coarse_dis[(size_t)q * nprobe + r] = ordered_to_float((uint32_t)(k >> 32));
    }
  }
  for (int r = c + tid; r < nprobe; r += CW_THREADS) {  // "not enough centroids": key -1
    keys[(size_t)q * nprobe + r] = -1;
    coarse_dis[(size_t)q * nprobe + r] = 3.402823466e38f;
  }
}

// ---------------------------------------------------------------------------------------------
// Register-resident variant of the row select for nlist <= 256 * 4 * V (16384 at V = 16): the row is read from
// HBM/L2 exactly once (V independent 16-byte loads per thread) and both passes run on registers.  Pass 1 is a
// plain float minimum (NaN -> +inf), tau0 = the largest of the G group minima as a VALUE (every column with a
// value <= tau0 survives, which is a superset of the key-based bound), pass 2 appends the ~G ln G survivors with
// one shared atomic each (rare), ranking as above.  The bisection fallback (survivors overflow the buffer,
// e.g. thousands of equal distances) runs on the 64-bit keys so that it terminates on ties.
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(CW_THREADS, 2) coarse_select_reg_kernel(const float *__restrict__ dist, int nlist,
                                                                       int nprobe, int G, int *__restrict__ keys,
                                                                       float *__restrict__ coarse_dis) {
  __shared__ u64 cand[CW_CAP];
  __shared__ float gminf[CW_THREADS];
  __shared__ float s_tauf;
  __shared__ int s_cnt;
  const int tid = threadIdx.x;
  const int q = blockIdx.x;
  const float *row = dist + (size_t)q * nlist;
  const float INF = __int_as_float(0x7f800000);
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; i++) {
    const int c = (i * CW_THREADS + tid) * 4;
    v[i] = c < nlist ? __ldg(reinterpret_cast<const float4 *>(row + c)) : make_float4(INF, INF, INF, INF);
  }
  float m = INF;
#pragma unroll
  for (int i = 0; i < V; i++) {
    v[i].x = v[i].x == v[i].x ? v[i].x : INF;  // NaN never selected (cw_key gives it the neutral key)
    v[i].y = v[i].y == v[i].y ? v[i].y : INF;
    v[i].z = v[i].z == v[i].z ? v[i].z : INF;
    v[i].w = v[i].w == v[i].w ? v[i].w : INF;
    m = fminf(m, fminf(fminf(v[i].x, v[i].y), fminf(v[i].z, v[i].w)));
  }
  gminf[tid] = m;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  if (tid < 32) {  // fold the 256 thread minima into G group minima, tau0 = their maximum
    float t = -INF;
    for (int g = tid; g < G; g += 32) {
      float gm = INF;
      for (int j = g; j < CW_THREADS; j += G) gm = fminf(gm, gminf[j]);
      t = fmaxf(t, gm);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) t = fmaxf(t, __shfl_xor_sync(GB_FULL, t, o));
    if (tid == 0) s_tauf = t;
  }
  __syncthreads();
  const float tauf = s_tauf;  // +inf: some group has no usable value, every finite column survives
  auto visit = [&](float x, int c) {
    if (x <= tauf && x < INF) {
      const int slot = atomicAdd(&s_cnt, 1);
      if (slot < CW_CAP) cand[slot] = ((u64)float_to_ordered(x) << 32) | (uint32_t)c;
    }
  };
#pragma unroll
  for (int i = 0; i < V; i++) {
    const int c = (i * CW_THREADS + tid) * 4;
    visit(v[i].x, c), visit(v[i].y, c + 1), visit(v[i].z, c + 2), visit(v[i].w, c + 3);
  }
  __syncthreads();
  int cnt = s_cnt;
  if (cnt > CW_CAP) {  // slow path: bisection on the key space
    u64 lo_b = 0, hi_b = ((u64)float_to_ordered(tauf) << 32) | 0xffffffffu, tau = hi_b;
    for (int attempt = 0; attempt < 130; attempt++) {
      tau = lo_b + (hi_b - lo_b) / 2;
      __syncthreads();
      if (tid == 0) s_cnt = 0;
      __syncthreads();
      for (int c = tid; c < nlist; c += CW_THREADS) {  // rare path: re-read the row (L2) instead of unrolling over registers
        const u64 key = cw_key(row[c], c);
        if (key != GB_KEY_MAX && key <= tau) {
          const int slot = atomicAdd(&s_cnt, 1);
          if (slot < CW_CAP) cand[slot] = key;
        }
      }
      __syncthreads();
      cnt = s_cnt;
      if (cnt <= CW_CAP && cnt >= nprobe) break;
      if (cnt > CW_CAP) hi_b = tau;
      else lo_b = tau + 1;
    }
  }
  const int c = min(cnt, CW_CAP);
  for (int i = tid; i < c; i += CW_THREADS) {  // rank by counting; survivors are unique keys
    const u64 k = cand[i];
    int r = 0;
    for (int j = 0; j < c; j++) r += cand[j] < k;
    if (r < nprobe) {
      keys[(size_t)q * nprobe + r] = (int)(uint32_t)k;
      coarse_dis[(size_t)q * nprobe + r] = ordered_to_float((uint32_t)(k >> 32));
    }
  }
  for (int r = c + tid; r < nprobe; r += CW_THREADS) {  // "not enough centroids": key -1
    keys[(size_t)q * nprobe + r] = -1;
    coarse_dis[(size_t)q * nprobe + r] = 3.402823466e38f;
  }
}

// ---------------------------------------------------------------------------------------------
// Select that starts from the GEMM's chunk minima (tc_gemm.cu epilogue: cmin[q][c] = min of dist[q][32c .. 32c+31]).
//   1. 128 threads fold the row's chunk minima into 128 group minima (consecutive chunks per thread);
//   2. tau = the nprobe-th smallest group minimum: nprobe distinct groups hold an element <= tau, so the nprobe-th
//      smallest distance of the row is <= tau;
//   3. only chunks whose minimum is <= tau can hold a survivor (about nprobe .. 2 nprobe of them): their 32 distances
//      are read (128 B per chunk, coalesced), survivors <= tau are ranked by counting — ascending distance, ties by
//      centroid id, as every other select here.
// The row itself is touched for ~nprobe x 128 B instead of nlist x 4 B (64 KB at nlist = 16384).  Rows the bound cannot
// handle (fewer than nprobe finite group minima, thousands of equal distances) fall back to a bisection over the key
// space on the full row: slow, rare, exact.
// ---------------------------------------------------------------------------------------------
constexpr int CM_THREADS = 128;
constexpr int CM_CAP = 1024;     // survivors
constexpr int CM_CHUNKS = 512;   // candidate chunks

__global__ void __launch_bounds__(CM_THREADS) coarse_select_cmin_kernel(const float *__restrict__ dist,
                                                                        const float *__restrict__ cmin, int cmin_pitch,
                                                                        int nlist, int nprobe, int *__restrict__ keys,
                                                                        float *__restrict__ coarse_dis) {
  __shared__ u64 cand[CM_CAP];
  __shared__ int chunk_list[CM_CHUNKS];
  __shared__ __align__(16) float gm_s[CM_THREADS];
  __shared__ float s_tauf;
  __shared__ int s_cnt, s_nch;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = blockIdx.x;
  const float *row = dist + (size_t)q * nlist;
  const float *cm = cmin + (size_t)q * cmin_pitch;
  const float INF = __int_as_float(0x7f800000);
  // chunks per thread, a multiple of 4 so that a thread's share is whole float4s (cmin_pitch is a multiple of 4)
  const int E = (((cmin_pitch + CM_THREADS - 1) / CM_THREADS) + 3) & ~3;
  const int c_lo = tid * E, c_hi = min(c_lo + E, cmin_pitch);
  float gm = INF;
  // a thread's chunk minima stay in registers when there are at most 16 of them (nlist <= 65536): the candidate pass
  // below needs them again and must not wait for a second round trip
  constexpr int KEEP = 16;
  float kept[KEEP];
#pragma unroll
  for (int i = 0; i < KEEP; i += 4) {
    const int c = c_lo + i;
    float4 v = make_float4(INF, INF, INF, INF);
    if (i < E && c < c_hi) v = __ldg(reinterpret_cast<const float4 *>(cm + c));
    kept[i] = v.x, kept[i + 1] = v.y, kept[i + 2] = v.z, kept[i + 3] = v.w;
    gm = fminf(gm, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));  // fminf drops NaN (a chunk of nothing but NaN)
  }
  for (int c = c_lo + KEEP; c < c_hi; c += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(cm + c));
    gm = fminf(gm, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
  }
  gm = gm < INF ? gm : INF;
  gm_s[tid] = gm;
  if (tid == 0) {
    s_tauf = INF;
    s_cnt = 0;
    s_nch = 0;
  }
  __syncthreads();
  {
    int r = 0;
#pragma unroll 8
    for (int j = 0; j < CM_THREADS; j += 4) {  // broadcast 128-bit reads
      const float4 o = *reinterpret_cast<const float4 *>(gm_s + j);
      r += (o.x < gm) || (o.x == gm && j < tid);
      r += (o.y < gm) || (o.y == gm && j + 1 < tid);
      r += (o.z < gm) || (o.z == gm && j + 2 < tid);
      r += (o.w < gm) || (o.w == gm && j + 3 < tid);
    }
    if (r == nprobe - 1 && gm < INF) s_tauf = gm;
  }
  __syncthreads();
  const float tauf = s_tauf;
  bool slow = !(tauf < INF);
  if (!slow) {
#pragma unroll
    for (int i = 0; i < KEEP; i++) {
      if (kept[i] <= tauf) {  // +inf beyond the thread's share
        const int slot = atomicAdd(&s_nch, 1);
        if (slot < CM_CHUNKS) chunk_list[slot] = c_lo + i;
      }
    }
    for (int c = c_lo + KEEP; c < c_hi; c++) {
      if (cm[c] <= tauf) {
        const int slot = atomicAdd(&s_nch, 1);
        if (slot < CM_CHUNKS) chunk_list[slot] = c;
      }
    }
    __syncthreads();
    const int nch = s_nch;
    if (nch > CM_CHUNKS) {
      slow = true;
    } else {
      for (int i = warp; i < nch; i += CM_THREADS / 32) {
        const int c = chunk_list[i] * 32 + lane;
        const float x = c < nlist ? __ldg(row + c) : INF;
        const bool pass = x <= tauf;
        const unsigned m = __ballot_sync(GB_FULL, pass);
        if (m) {
          int base = 0;
          const int leader = __ffs(m) - 1;
          if (lane == leader) base = atomicAdd(&s_cnt, __popc(m));
          base = __shfl_sync(GB_FULL, base, leader);
          const int slot = base + __popc(m & ((1u << lane) - 1u));
          if (pass && slot < CM_CAP) cand[slot] = ((u64)float_to_ordered(x) << 32) | (uint32_t)c;
        }
      }
      __syncthreads();
      if (s_cnt > CM_CAP) slow = true;
    }
  }
  int cnt = s_cnt;
  if (slow) {  // bisection on the key space over the full row
    u64 lo_b = 0, hi_b = ((u64)float_to_ordered(tauf) << 32) | 0xffffffffu, tau = hi_b;
    for (int attempt = 0; attempt < 132; attempt++) {
      __syncthreads();
      if (tid == 0) s_cnt = 0;
      __syncthreads();
      for (int c = tid; c < nlist; c += CM_THREADS) {
        const u64 key = cw_key(row[c], c);
        if (key != GB_KEY_MAX && key <= tau) {
          const int slot = atomicAdd(&s_cnt, 1);
          if (slot < CM_CAP) cand[slot] = key;
        }
      }
      __syncthreads();
      cnt = s_cnt;
      if (cnt <= CM_CAP && (cnt >= nprobe || attempt == 0)) break;  // attempt 0 sees everything at or below tau0
      if (cnt > CM_CAP) hi_b = tau;
      else lo_b = tau + 1;
      tau = lo_b + (hi_b - lo_b) / 2;
    }
  }
  const int c = min(cnt, CM_CAP);
  for (int i = tid; i < c; i += CM_THREADS) {  // rank by counting; survivors are unique keys
    const u64 k = cand[i];
    int r = 0;
    for (int j = 0; j < c; j++) r += cand[j] < k;
    if (r < nprobe) {
      keys[(size_t)q * nprobe + r] = (int)(uint32_t)k;
      coarse_dis[(size_t)q * nprobe + r] = ordered_to_float((uint32_t)(k >> 32));
    }
  }
  for (int r = c + tid; r < nprobe; r += CM_THREADS) {  // "not enough centroids": key -1
    keys[(size_t)q * nprobe + r] = -1;
    coarse_dis[(size_t)q * nprobe + r] = 3.402823466e38f;
  }
}

// at least one chunk per group and the bound needs nprobe <= groups
bool coarse_select_cmin_usable(int nlist, int nprobe) {
  return nprobe <= CM_THREADS && nprobe >= 1 && (nlist + 31) / 32 >= CM_THREADS && nlist >= nprobe;
}

cudaError_t launch_coarse_select_cmin(const float *dist, const float *cmin, int cmin_pitch, int n, int nlist, int nprobe,
                                      int *keys, float *coarse_dis, cudaStream_t st) {
  if (!coarse_select_cmin_usable(nlist, nprobe) || (cmin_pitch & 3)) return cudaErrorInvalidValue;
  coarse_select_cmin_kernel<<<n, CM_THREADS, 0, st>>>(dist, cmin, cmin_pitch, nlist, nprobe, keys, coarse_dis);
  return cudaGetLastError();
}

cudaError_t launch_coarse_select(const float *dist, int n, int nlist, int nprobe, int *keys, float *coarse_dis,
                                 cudaStream_t st) {
  if (nprobe <= 128) {
    const int G = nprobe <= 32 ? 32 : (nprobe <= 64 ? 64 : 128);
    const bool aligned = (nlist & 3) == 0 && ((uintptr_t)dist & 15) == 0;
    if (aligned && nlist <= CW_THREADS * 4 * 16 && nlist >= nprobe) {
      if (nlist <= CW_THREADS * 4 * 4)
        coarse_select_reg_kernel<4><<<n, CW_THREADS, 0, st>>>(dist, nlist, nprobe, G, keys, coarse_dis);
      else
        coarse_select_reg_kernel<16><<<n, CW_THREADS, 0, st>>>(dist, nlist, nprobe, G, keys, coarse_dis);
      return cudaGetLastError();
    }
    coarse_select_row_kernel<<<n, CW_THREADS, 0, st>>>(dist, nlist, nprobe, G, keys, coarse_dis);
    return cudaGetLastError();
  }
  int need = nprobe + CS_THREADS * CS_PER_ROUND;
  int cap = 1024;
  while (cap < need) cap <<= 1;
  if (cap < next_pow2(nprobe)) cap = next_pow2(nprobe);
  size_t smem = (size_t)cap * sizeof(u64) + (4 + 64) * sizeof(int);
  if (cap <= 4 * CS_THREADS) {
    coarse_select_kernel<4><<<n, CS_THREADS, smem, st>>>(dist, nlist, nprobe, cap, keys, coarse_dis);
  } else {
    cudaError_t e = ensure_dynamic_smem(coarse_select_kernel<16>, smem);
    if (e != cudaSuccess) return e;
    coarse_select_kernel<16><<<n, CS_THREADS, smem, st>>>(dist, nlist, nprobe, cap, keys, coarse_dis);
  }
  return cudaGetLastError();
}

}  // namespace gb
