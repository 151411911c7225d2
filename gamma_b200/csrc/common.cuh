// Common device helpers for the gamma_b200 search kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.h"

#define GB_WARP 32
#define GB_FULL 0xffffffffu

namespace gb {


// ---------------------------------------------------------------------------
// Orderable keys.
// A candidate is the 64-bit key  (ordered_distance << 32) | seq, where
//   ordered_distance : u32 such that SMALLER key == BETTER candidate
//                      (L2: ascending distance; InnerProduct: descending score)
//   seq              : scan order of the posting, (probe_rank << 21) | position_in_list,
//                      so that among equal distances the FIRST SCANNED posting wins —
//                      the reference's strict `C::cmp(heap[0], dis)` test rejects a later
//                      equal candidate (gamma_index_ivfpq.h:363-368).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending float order == ascending u32
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}
// L2 keeps the smallest distances, IP the largest scores.
template <bool IP>
__device__ __forceinline__ uint32_t dist_to_key32(float d) {
  uint32_t o = float_to_ordered(d);
  return IP ? ~o : o;
}
template <bool IP>
__device__ __forceinline__ float key32_to_dist(uint32_t k) {
  return ordered_to_float(IP ? ~k : k);
}
__device__ __forceinline__ float key32_to_dist(uint32_t k, bool ip) {
  return ordered_to_float(ip ? ~k : k);
}
__device__ __forceinline__ uint32_t dist_to_key32(float d, bool ip) {
  uint32_t o = float_to_ordered(d);
  return ip ? ~o : o;
}

#define GB_KEY_MAX 0xffffffffffffffffull

// validity bitmap: bit = 1 <=> doc may be returned (NOT deleted AND passes all range filters)
__device__ __forceinline__ bool bitmap_test(const uint32_t *__restrict__ bm, int id) {
  return (__ldg(bm + (id >> 5)) >> (id & 31)) & 1u;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int ldg_nc_s32(const void *p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_nc_f32(const void *p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------
// BlockTopR — streaming "keep the R smallest 64-bit keys" for one CTA.
//
//   * append(): a candidate whose key is below the current threshold tau is pushed
//     into an UNSORTED shared-memory buffer (one warp-aggregated atomicAdd);
//   * prune(): when the buffer is (nearly) full the whole CTA finds the R-th smallest key
//     with a register-resident radix select (see prune_impl), compacts the survivors to
//     the front and tightens tau.
//   No sorting network runs during the scan; only the final <=R survivors are sorted
//   where an order is required.
// Keys are unique (seq is unique per posting), so a separating threshold exists.
// All methods marked "collective" must be called by every thread of the CTA.
// ---------------------------------------------------------------------------
struct BlockTopR {
  u64 *buf;            // [cap] shared
  int *cnt;            // shared
  u64 *tau;            // shared: current admission threshold (exclusive)
  int *warp_part;      // [2][32] shared scratch for counts
  int cap;
  int R;
  // optional float image of tau's distance word (the v3 scan pre-tests candidates with one FSETP against it):
  // L2: admit only dis <= *tau_f, InnerProduct: dis >= *tau_f.  nullptr = not maintained.
  float *tau_f = nullptr;
  int is_ip = 0;

  __device__ __forceinline__ void init_collective() {
    if (threadIdx.x == 0) {
      *cnt = 0;
      *tau = GB_KEY_MAX;
    }
    __syncthreads();
  }

  __device__ __forceinline__ u64 threshold() const { return *((volatile u64 *)tau); }

  // warp-collective (all 32 lanes call it, `pass` may differ per lane)
  __device__ __forceinline__ void append_warp(bool pass, u64 key) {
    unsigned m = __ballot_sync(GB_FULL, pass);
    if (m == 0) return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(cnt, __popc(m));
    base = __shfl_sync(GB_FULL, base, leader);
    if (pass) {
      int slot = base + __popc(m & ((1u << lane) - 1u));
      if (slot < cap) buf[slot] = key;  // cap is never exceeded by construction; guard anyway
    }
  }

  // collective: keep only the R smallest keys (no-op when <= R are buffered), tighten tau.
  //
  // Radix select on registers: every thread pulls its share of the buffer into registers (the
  // buffer then doubles as scratch), the CTA narrows [min,max] of the distance word by a 256-bin
  // shared-memory histogram per pass (power-of-two bins => exact integer boundaries) until the bin
  // holding the R-th key has <= 64 members, ranks those 64 by brute force, and compacts.  Equal
  // distance words fall through to a second stage on the scan-order word.  About a dozen barriers
  // and ~2 passes in practice, versus 64 barrier-separated bisection steps.  PER = keys held per thread: cap <= PER * blockDim.x
  // (the launchers pick 4 when the buffer has <= 4 keys per thread, else 16).
  // EXACT = false (in-loop prunes of the persistent scan): stop after the first histogram pass whose bin boundary leaves
  // at most R + (cap - R) / 4 survivors — every key up to the upper edge of the bin that holds the R-th one is kept, tau
  // becomes that edge.  The R best are all among the survivors, the threshold is only slightly looser than the exact one,
  // and the brute-force ranking and its barriers are skipped; the query's final select runs with EXACT = true.
  template <int PER, bool EXACT = true>
  __device__ __forceinline__ void prune_collective() {
    __syncthreads();
    const int n = min(*((volatile int *)cnt), cap);
    if (n <= R) {  // nothing to drop (uniform decision)
      __syncthreads();
      return;
    }
    const int tid = threadIdx.x, bd = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = (bd + 31) >> 5;
    u64 k[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
      int i = tid + j * bd;
      k[j] = i < n ? buf[i] : GB_KEY_MAX;
    }
    __syncthreads();  // buf is scratch from here until the compaction
    int *hist = reinterpret_cast<int *>(buf);          // [256]
    volatile int *var = hist + 256;                    // [16]: 0 bin, 1 below, 2 in-bin, 4 ncand, 6..7 t*
    u64 *cand = buf + 160;                             // [64] at byte 1280
    int stage = 0;      // 0: distance word, 1: scan-order word among keys whose distance word == fixed_hi
    uint32_t fixed_hi = 0, lo = 0, hi = 0;
    int base = 0;       // keys known to be strictly below the current range
    u64 tstar = 0;
    bool resolved = false;

    auto in_scope = [&](u64 key) -> bool {
      return key != GB_KEY_MAX && (stage == 0 || (uint32_t)(key >> 32) == fixed_hi);
    };
    auto word = [&](u64 key) -> uint32_t { return stage == 0 ? (uint32_t)(key >> 32) : (uint32_t)key; };
    auto minmax = [&]() {  // collective: lo/hi = min/max word over keys in scope
      uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
      for (int j = 0; j < PER; j++)
        if (in_scope(k[j])) {
          uint32_t w = word(k[j]);
          mn = min(mn, w);
          mx = max(mx, w);
        }
      mn = __reduce_min_sync(GB_FULL, mn);
      mx = __reduce_max_sync(GB_FULL, mx);
      if (lane == 0) {
        hist[wid] = (int)mn;
        hist[32 + wid] = (int)mx;
      }
      __syncthreads();
      mn = 0xffffffffu;
      mx = 0u;
      for (int w = 0; w < nw; w++) {
        mn = min(mn, (uint32_t)hist[w]);
        mx = max(mx, (uint32_t)hist[32 + w]);
      }
      lo = mn;
      hi = mx;
      __syncthreads();
    };

    minmax();
    for (;;) {
      if (lo == hi) {
        if (stage == 0) {  // every remaining key has the same distance word: decide on scan order
          stage = 1;
          fixed_hi = lo;
          minmax();
          continue;
        }
        tstar = ((u64)fixed_hi << 32) | lo;  // keys are unique: a single key is left
        resolved = true;
        break;
      }
      const uint32_t range = hi - lo;
      const int shift = max(0, (32 - __clz(range)) - 8);  // (range >> shift) < 256
      if (tid < 256) hist[tid] = 0;
      if (tid == 0) var[4] = 0;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < PER; j++)
        if (in_scope(k[j])) {
          uint32_t w = word(k[j]);
          if (w >= lo && w <= hi) atomicAdd(&hist[(w - lo) >> shift], 1);
        }
      __syncthreads();
      if (wid == 0) {  // locate the bin that holds rank `need` (1-based) inside the range
        int c[8], s = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
          c[t] = hist[lane * 8 + t];
          s += c[t];
        }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int v = __shfl_up_sync(GB_FULL, incl, o);
          if (lane >= o) incl += v;
        }
        const int excl = incl - s, need = R - base;
        if (excl < need && need <= incl) {
          int run = excl;
#pragma unroll
          for (int t = 0; t < 8; t++) {
            if (need > run && need <= run + c[t]) {
              var[0] = lane * 8 + t;
              var[1] = run;
              var[2] = c[t];
            }
            run += c[t];
          }
        }
      }
      __syncthreads();
      const int b = var[0], below = var[1], cb = var[2];
      base += below;
      const uint32_t nlo = lo + ((uint32_t)b << shift);
      const uint32_t nhi = shift == 0 ? nlo : min(hi, nlo + ((1u << shift) - 1u));
      lo = nlo;
      hi = nhi;
      __syncthreads();  // scratch is reused by the next pass / the gather
      if (!EXACT && stage == 0 && base + cb <= R + ((cap - R) >> 2)) {  // uniform: good enough for an in-loop prune
        tstar = ((u64)hi << 32) | 0xffffffffu;
        resolved = true;
        break;
      }
      if (cb <= 64) break;
    }
    if (!resolved) {
      // gather the <= 64 keys of the final range and rank them by brute force
#pragma unroll
      for (int j = 0; j < PER; j++)
        if (in_scope(k[j])) {
          uint32_t w = word(k[j]);
          if (w >= lo && w <= hi) {
            int sidx = atomicAdd((int *)&var[4], 1);
            if (sidx < 64) cand[sidx] = k[j];
          }
        }
      __syncthreads();
      const int nc = min((int)var[4], 64), need = R - base;
      if (wid == 0) {
        for (int i = lane; i < nc; i += 32) {
          const u64 c = cand[i];
          int r = 0;
          for (int j = 0; j < nc; j++) r += cand[j] < c;
          if (r == need - 1) *reinterpret_cast<volatile u64 *>(var + 6) = c;
        }
      }
      __syncthreads();
      tstar = *reinterpret_cast<volatile u64 *>(var + 6);
    }
    // survivors: key <= t*  (exactly R of them)
    unsigned keep = 0;
#pragma unroll
    for (int j = 0; j < PER; j++)
      if (k[j] != GB_KEY_MAX && k[j] <= tstar) keep |= (1u << j);
    __syncthreads();
    if (tid == 0) *cnt = 0;
    __syncthreads();
    if (keep) {
      int o = atomicAdd(cnt, __popc(keep));
#pragma unroll
      for (int j = 0; j < PER; j++)
        if ((keep >> j) & 1u) buf[o++] = k[j];
    }
    if (tid == 0) {
      *tau = tstar + 1;  // admit only strictly better than the R-th
      if (tau_f) *tau_f = key32_to_dist((uint32_t)((tstar + 1) >> 32), is_ip != 0);
    }
    __syncthreads();
  }

};

// ---------------------------------------------------------------------------
// Exact L2^2 / inner product in the summation order of faiss' AVX kernels (utils/distances_simd.cpp:366-431, as
// compiled for the reference): 8 strided partial sums (mul then add, unfused), hi + lo halves, fused 4-wide and masked
// tails, two horizontal adds — re-ranked and FLAT distances are bit-identical to the CPU engine's.
// ---------------------------------------------------------------------------
// 8 consecutive lanes (an "octet") cooperate on one candidate; returns the result in every lane of the octet
// YS: y points into shared memory (rows staged by the caller) instead of global memory
template <bool IP, bool YS = false>
__device__ __forceinline__ float exact_distance_octet(const float *__restrict__ q, const float *__restrict__ y,
                                                      int d, int sub /*0..7*/) {
  float s = 0.f;
  int d8 = d & ~7;
  // 16 strided elements of the row per lane are requested before the first one is used (profiles/r02_rerank: with
  // two loads in flight 41 % of the kernel's warp time sat on the first FADD after each load); the arithmetic and its
  // order are unchanged
  constexpr int U = 16;
  for (int i0 = sub; i0 < d8; i0 += 8 * U) {
    float yv[U];
#pragma unroll
    for (int u = 0; u < U; u++) yv[u] = (i0 + 8 * u < d8) ? (YS ? y[i0 + 8 * u] : __ldg(y + i0 + 8 * u)) : 0.f;
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (i0 + 8 * u < d8) {
        const float a = q[i0 + 8 * u], b = yv[u];
        if (IP) {
          s = __fadd_rn(s, __fmul_rn(a, b));
        } else {
          float t = __fsub_rn(a, b);
          s = __fadd_rn(s, __fmul_rn(t, t));
        }
      }
    }
  }
  // msum2 = hi + lo
  float other = __shfl_down_sync(GB_FULL, s, 4, 8);
  float t4 = __fadd_rn(other, s);  // valid in sub 0..3
  int rem = d - d8;
  if (rem >= 4) {
    if (sub < 4) {
      float a = q[d8 + sub], b = YS ? y[d8 + sub] : __ldg(y + d8 + sub);
      t4 = IP ? __fmaf_rn(a, b, t4) : __fmaf_rn(__fsub_rn(a, b), __fsub_rn(a, b), t4);
    }
    d8 += 4;
    rem -= 4;
  }
  if (rem > 0) {
    if (sub < rem) {
      float a = q[d8 + sub], b = YS ? y[d8 + sub] : __ldg(y + d8 + sub);
      t4 = IP ? __fmaf_rn(a, b, t4) : __fmaf_rn(__fsub_rn(a, b), __fsub_rn(a, b), t4);
    }
  }
  // hadd, hadd: (t0 + t1) + (t2 + t3)
  float n1 = __shfl_xor_sync(GB_FULL, t4, 1, 8);
  float p = __fadd_rn(t4, n1);  // lanes 0,1: t0+t1 ; lanes 2,3: t2+t3
  float n2 = __shfl_xor_sync(GB_FULL, p, 2, 8);
  float r = __fadd_rn(p, n2);
  return __shfl_sync(GB_FULL, r, 0, 8);
}

// In-place bitonic sort (ascending) of n = power of two u64 keys in shared memory. collective.
__device__ __forceinline__ void block_bitonic_sort(u64 *a, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          u64 x = a[i], y = a[ixj];
          bool up = ((i & k) == 0);
          if ((x > y) == up) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
    }
  }
  __syncthreads();
}

__host__ __device__ inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace gb
