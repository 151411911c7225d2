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
//   * prune(): when the buffer cannot take another round, the whole CTA finds the
//     R-th smallest key by MSB-first bisection over the 64-bit key space
//     (count-below per step: register compares + one warp REDUX + one barrier —
//     almost no shared-memory/LSU traffic, which the ADC lookups need), compacts
//     the survivors to the front and tightens tau.
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

  // collective: number of buffered keys strictly below t
  __device__ __forceinline__ int count_below(u64 t, int n, int parity) {
    int c = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c += (buf[i] < t);
    c = __reduce_add_sync(GB_FULL, c);
    int nw = (blockDim.x + 31) >> 5;
    int *part = warp_part + parity * 32;
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
    __syncthreads();
    int tot = 0;
    for (int w = 0; w < nw; w++) tot += part[w];
    return tot;
  }

  // collective: keep only the `keep` smallest keys (keep = min(R, count)), set tau.
  __device__ void prune_collective() {
    __syncthreads();
    int n = min(*((volatile int *)cnt), cap);
    if (n <= R) {  // nothing to drop (uniform decision)
      __syncthreads();
      return;
    }
    // Find the smallest t with count(key < t) >= R, bit by bit: t = prefix with the
    // undecided low bits zero; invariant count(key < lo) < R.  After 64 steps lo is the
    // R-th smallest key itself, so count(key <= lo) == R because keys are unique.
    u64 lo = 0;
    int parity = 0;
    for (int bit = 63; bit >= 0; --bit) {
      u64 cand = lo | (1ull << bit);
      int c = count_below(cand, n, parity);
      parity ^= 1;
      if (c < R) {
        lo = cand;  // R-th smallest is >= cand
      } else if (c == R) {
        // cand separates exactly R keys: done early.
        lo = cand - 1;  // keys <= lo are the survivors
        break;
      }
    }
    // survivors: key <= lo  (exactly R of them when the loop ran to the end: lo == R-th key)
    // compact in two phases through registers (buffer is read fully before being rewritten)
    const int PER = 16;  // cap <= PER * blockDim.x is guaranteed by the launcher
    u64 mine[PER];
    unsigned keep = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) {
      int i = threadIdx.x + j * blockDim.x;
      mine[j] = 0;
      if (i < n) {
        mine[j] = buf[i];
        if (mine[j] <= lo) keep |= (1u << j);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) *cnt = 0;
    __syncthreads();
    if (keep) {
      int o = atomicAdd(cnt, __popc(keep));
#pragma unroll
      for (int j = 0; j < PER; j++)
        if ((keep >> j) & 1u) buf[o++] = mine[j];
    }
    if (threadIdx.x == 0) *tau = lo + 1;  // admit only strictly better than the R-th
    __syncthreads();
  }
};

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
