// K5 — the encode half of GammaIVFPQIndex::Add on the device (index/impl/gamma_index_ivfpq.cc:424-476):
//   quantizer->assign            -> the coarse stage with nprobe = 1 (tc_gemm.cu + coarse.cu, launched by capi.cu)
//   compute_residuals            -> r = x - centroid[key]                       (one rounding, as Index::compute_residual)
//   pq.compute_codes(r)          -> per sub-quantiser argmin_c |r_m - cb[m][c]|^2 (faiss ProductQuantizer.cpp:320-347)
// The argmin must pick the same centroid as the CPU engine, so the sub-distances reproduce the arithmetic of faiss'
// fvec_L2sqr_ny as compiled for the reference (utils/distances_simd.cpp:205-317, built with -mavx2 -mfma):
//   dsub 1 : (x-y)^2
//   dsub 2 : d0^2 + d1^2                               (mul, mul, hadd)
//   dsub 4 : (d0^2 + d1^2) + (d2^2 + d3^2)             (mul x4, hadd, hadd)
//   dsub 8 : a_i = fma(d_i, d_i, d_{i+4}^2), (a0 + a1) + (a2 + a3)
//   dsub 12: a_i = fma(d_{i+8}, d_{i+8}, fma(d_i, d_i, d_{i+4}^2)), (a0 + a1) + (a2 + a3)
//   other dsub < 16: fvec_L2sqr's AVX order (8 strided partial sums, mul then add; hi + lo; fused 4-wide and masked tails;
//                    two horizontal adds) — the same order rerank.cu reproduces for whole vectors.
//   dsub >= 16: faiss switches to distance tables through sgemm, whose summation order is the BLAS library's; there the
//               codes can differ on floating-point near-ties only.
// Ties keep the lowest centroid index (strict `dis < mindis`).
#include "common.cuh"
#include "kernels.h"

namespace gb {

template <int DSUB>
__device__ __forceinline__ float sub_l2_exact(const float *x, const float *y) {
  float d[DSUB];
#pragma unroll
  for (int i = 0; i < DSUB; i++) d[i] = __fsub_rn(x[i], y[i]);
  if (DSUB == 1) return __fmul_rn(d[0], d[0]);
  if (DSUB == 2) return __fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1]));
  float a[4];
  if (DSUB == 4) {
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = __fmul_rn(d[i], d[i]);
  } else if (DSUB == 8) {
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = __fmaf_rn(d[i], d[i], __fmul_rn(d[4 + i], d[4 + i]));
  } else {  // 12
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = __fmaf_rn(d[8 + i], d[8 + i], __fmaf_rn(d[i], d[i], __fmul_rn(d[4 + i], d[4 + i])));
  }
  return __fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3]));
}

// fvec_L2sqr (AVX) order for any dsub, one thread
__device__ __forceinline__ float sub_l2_avx_order(const float *x, const float *y, int dsub) {
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int d8 = dsub & ~7;
  for (int i = 0; i < d8; i += 8)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float t = __fsub_rn(x[i + j], y[i + j]);
      s[j] = __fadd_rn(s[j], __fmul_rn(t, t));
    }
  float t4[4];
#pragma unroll
  for (int j = 0; j < 4; j++) t4[j] = __fadd_rn(s[j + 4], s[j]);
  int rem = dsub - d8;
  if (rem >= 4) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float t = __fsub_rn(x[d8 + j], y[d8 + j]);
      t4[j] = __fmaf_rn(t, t, t4[j]);
    }
    d8 += 4;
    rem -= 4;
  }
  for (int j = 0; j < rem; j++) {
    const float t = __fsub_rn(x[d8 + j], y[d8 + j]);
    t4[j] = __fmaf_rn(t, t, t4[j]);
  }
  return __fadd_rn(__fadd_rn(t4[0], t4[1]), __fadd_rn(t4[2], t4[3]));
}

// One warp = 32 vectors x one sub-quantiser at a time (every lane reads the same centroid: broadcast loads, L1 resident
// — a sub-quantiser's codebook is 256 x dsub floats); the warps of a CTA share the 32 vectors and split the M
// sub-quantisers.
template <int DSUB>
__global__ void __launch_bounds__(256) pq_encode_kernel(const float *__restrict__ x, int x_stride,
                                                        const int *__restrict__ keys,
                                                        const float *__restrict__ centroids,
                                                        const float *__restrict__ pq, long long n, int d, int M, int dsub_rt,
                                                        int by_residual, uint8_t *__restrict__ codes) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long v = (long long)blockIdx.x * 32 + lane;
  const bool have = v < n;
  const int dsub = DSUB > 0 ? DSUB : dsub_rt;
  const int key = have ? keys[v] : 0;
  const float *xr = x + (size_t)(have ? v : 0) * x_stride;
  const float *cr = centroids + (size_t)(key < 0 ? 0 : key) * d;
  for (int m = warp; m < M; m += nw) {
    constexpr int DS = (DSUB > 0) ? DSUB : 16;
    float r[DS];
    // residual of this vector's m-th slice; columns beyond the stored width are zero (ConvertVectorDim padding)
#pragma unroll
    for (int i = 0; i < DS; i++) {
      if (i < dsub) {
        const int col = m * dsub + i;
        const float xv = col < x_stride ? xr[col] : 0.f;
        r[i] = by_residual ? __fsub_rn(xv, cr[col]) : xv;
      } else {
        r[i] = 0.f;
      }
    }
    const float *cb = pq + (size_t)m * 256 * dsub;
    float best = 1e20f;
    int best_c = 0;
#pragma unroll 4
    for (int c = 0; c < 256; c++) {
      float y[DS];
#pragma unroll
      for (int i = 0; i < DS; i++) y[i] = i < dsub ? __ldg(cb + c * dsub + i) : 0.f;
      constexpr int DE = (DSUB > 0) ? DSUB : 1;
      const float dis = (DSUB > 0) ? sub_l2_exact<DE>(r, y) : sub_l2_avx_order(r, y, dsub);
      if (dis < best) {
        best = dis;
        best_c = c;
      }
    }
    if (have) codes[(size_t)v * M + m] = (uint8_t)best_c;
  }
}

// any dsub > 16: same arithmetic, the slice streamed from memory instead of held in registers
__global__ void __launch_bounds__(256) pq_encode_wide_kernel(const float *__restrict__ x, int x_stride,
                                                             const int *__restrict__ keys,
                                                             const float *__restrict__ centroids,
                                                             const float *__restrict__ pq, long long n, int d, int M, int dsub,
                                                             int by_residual, uint8_t *__restrict__ codes) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long v = (long long)blockIdx.x * 32 + lane;
  if (v >= n) return;
  const int key = keys[v];
  const float *xr = x + (size_t)v * x_stride;
  const float *cr = centroids + (size_t)(key < 0 ? 0 : key) * d;
  for (int m = warp; m < M; m += nw) {
    const float *cb = pq + (size_t)m * 256 * dsub;
    float best = 1e20f;
    int best_c = 0;
    for (int c = 0; c < 256; c++) {
      // fvec_L2sqr order with the 8 partial sums kept per lane
      float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float *y = cb + (size_t)c * dsub;
      int d8 = dsub & ~7;
      auto rv = [&](int i) {
        const int col = m * dsub + i;
        const float xv = col < x_stride ? xr[col] : 0.f;
        return by_residual ? __fsub_rn(xv, cr[col]) : xv;
      };
      for (int i = 0; i < d8; i += 8)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float t = __fsub_rn(rv(i + j), __ldg(y + i + j));
          s[j] = __fadd_rn(s[j], __fmul_rn(t, t));
        }
      float t4[4];
#pragma unroll
      for (int j = 0; j < 4; j++) t4[j] = __fadd_rn(s[j + 4], s[j]);
      int rem = dsub - d8;
      if (rem >= 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float t = __fsub_rn(rv(d8 + j), __ldg(y + d8 + j));
          t4[j] = __fmaf_rn(t, t, t4[j]);
        }
        d8 += 4;
        rem -= 4;
      }
      for (int j = 0; j < rem; j++) {
        const float t = __fsub_rn(rv(d8 + j), __ldg(y + d8 + j));
        t4[j] = __fmaf_rn(t, t, t4[j]);
      }
      const float dis = __fadd_rn(__fadd_rn(t4[0], t4[1]), __fadd_rn(t4[2], t4[3]));
      if (dis < best) {
        best = dis;
        best_c = c;
      }
    }
    codes[(size_t)v * M + m] = (uint8_t)best_c;
  }
}

cudaError_t launch_pq_encode(const float *x, int x_stride, const int *keys, const float *centroids, const float *pq,
                             long long n, int d, int M, int dsub, int by_residual, uint8_t *codes, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n + 31) / 32);
  const int threads = 32 * (M >= 8 ? 8 : M);
#define GB_ENC(DS) pq_encode_kernel<DS><<<grid, threads, 0, st>>>(x, x_stride, keys, centroids, pq, n, d, M, dsub, by_residual, codes)
  switch (dsub) {
    case 1: GB_ENC(1); break;
    case 2: GB_ENC(2); break;
    case 4: GB_ENC(4); break;
    case 8: GB_ENC(8); break;
    case 12: GB_ENC(12); break;
    default:
      if (dsub <= 16) GB_ENC(0);
      else pq_encode_wide_kernel<<<grid, threads, 0, st>>>(x, x_stride, keys, centroids, pq, n, d, M, dsub, by_residual, codes);
  }
#undef GB_ENC
  return cudaGetLastError();
}

}  // namespace gb
