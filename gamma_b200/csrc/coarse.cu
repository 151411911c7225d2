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

cudaError_t launch_coarse_select(const float *dist, int n, int nlist, int nprobe, int *keys, float *coarse_dis,
                                 cudaStream_t st) {
  int need = nprobe + CS_THREADS * CS_PER_ROUND;
  int cap = 1024;
  while (cap < need) cap <<= 1;
  if (cap < next_pow2(nprobe)) cap = next_pow2(nprobe);
  size_t smem = (size_t)cap * sizeof(u64) + (4 + 64) * sizeof(int);
  if (cap <= 4 * CS_THREADS) {
    coarse_select_kernel<4><<<n, CS_THREADS, smem, st>>>(dist, nlist, nprobe, cap, keys, coarse_dis);
  } else {
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      cudaError_t e = cudaFuncSetAttribute(coarse_select_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured = smem;
    }
    coarse_select_kernel<16><<<n, CS_THREADS, smem, st>>>(dist, nlist, nprobe, cap, keys, coarse_dis);
  }
  return cudaGetLastError();
}

}  // namespace gb
