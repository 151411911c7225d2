// K1t / K4t — dense Q x Y^T contraction on the 5th-generation tensor cores (tcgen05 + TMEM + TMA),
// the only GEMM-shaped work on the path: the coarse quantiser (n x nlist x d) and the flat scan
// (n x N x d).  Replaces the sgemm inside faiss knn_L2sqr (faiss utils/distances.cpp:215-296) and the
// per-vector fvec_L2sqr / fvec_inner_product loop of GammaFLATIndex::Search (gamma_index_flat.cc:183-232)
// as the *distance producer*; selection and (for FLAT) the exact re-score stay in their own kernels.
//
// Precision: operands are fp32 in memory and the tensor core multiplies TF32 (it ignores the low 13
// mantissa bits).  To keep fp32-level accuracy (the probe set must equal the CPU engine's) each
// product is evaluated as  big*big + big*small + small*big  with small = x - tf32(x) precomputed
// ("3xTF32"): three tcgen05.mma per k-step into the same TMEM accumulator, relative error ~2^-21.
//
// Structure (persistent CTA per SM over 128 x 128 output tiles, K pipelined in 32-float slabs, 3 smem stages,
// 2 TMEM accumulator stages so the epilogue of one tile overlaps the mainloop of the next):
//   warp 0 lane 0 : TMA producer  — cp.async.bulk.tensor.2d (SWIZZLE_128B) of the 4 operand tiles
//   warp 1 lane 0 : MMA issuer    — tcgen05.mma.cta_group::1.kind::tf32, tcgen05.commit -> mbarriers
//   warps 2..9    : epilogue      — tcgen05.ld (TMEM -> registers), |q|^2 + |y|^2 - 2 acc (clamped at 0)
//                                   or acc (InnerProduct), 128-bit stores of the distance tile
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "kernels.h"

namespace gb {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;  // tile; TC_BK floats = one 128-byte swizzle row
constexpr int TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;      // 16 KB per operand tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;     // A_big, A_small, B_big, B_small
constexpr int TC_EPI_WARPS = 8;                       // two warps per TMEM lane quarter, 64 columns each
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;    // TMA warp + MMA warp + epilogue warps
constexpr int TC_TMEM_COLS = 128;                     // fp32 accumulator: 128 lanes x 128 columns
constexpr int TC_TR_PITCH = 20;                       // epilogue transpose strip: 16 columns + 4 pad (16-byte rows)
constexpr int TC_TR_FLOATS = 32 * TC_TR_PITCH + 64;   // per epilogue warp: the strip + |y|^2 of its 64 columns

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tTCW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra TCD_%=;\n\tbra TCW_%=;\n\tTCD_%=:\n\t}" ::"r"(tc_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          tc_smem_u32(dst)),
      "l"(map), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major operand tile, SWIZZLE_128B: rows at 128 B pitch, 8-row atoms 1024 B apart (SBO), LBO unused.
// Bit layout: cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start[0,14) LBO[16,30) SBO[32,46)
// version[46,48)=1 layout_type[61,64)=2.
__device__ __forceinline__ uint64_t tc_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format[4,6)=1 (F32) a_format[7,10)=2 (TF32) b_format[10,13)=2, both K-major,
// n_dim[17,23)=N>>3, m_dim[24,29)=M>>4
__device__ __forceinline__ uint32_t tc_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar))
               : "memory");
}

struct TcGemmParams {
  const float *a_norm;  // [M] or nullptr
  const float *b_norm;  // [N] or nullptr
  float *out;           // [M][ldo]
  int M, N, K, ldo;
  int l2;               // 1: out = an + bn - 2 acc clamped at 0 ; 0: out = acc
  float *cmin;          // optional (l2 only): [M][cmin_pitch] minimum of every 32-column chunk of out (NaN ignored,
  int cmin_pitch;       //   columns >= N count as +inf); the coarse select starts from these instead of the full row
};

__device__ __forceinline__ void tc_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

// Persistent: one CTA per SM walks the output tiles; the TMEM accumulator is double-buffered
// (2 x 128 columns) so the epilogue of tile i overlaps the TMA/MMA of tile i+1.
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_big, const __grid_constant__ CUtensorMap map_a_small,
                      const __grid_constant__ CUtensorMap map_b_big, const __grid_constant__ CUtensorMap map_b_small,
                      TcGemmParams P) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char *tiles = tc_smem;
  uint64_t *full = reinterpret_cast<uint64_t *>(tc_smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t *empty = full + TC_STAGES;
  uint64_t *tmem_full = empty + TC_STAGES;   // [2]
  uint64_t *tmem_empty = tmem_full + 2;      // [2]
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
  float *tr_all = reinterpret_cast<float *>(tmem_ptr + 4);  // TC_EPI_WARPS x TC_TR_FLOATS, 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (P.M + TC_BM - 1) / TC_BM, tiles_n = (P.N + TC_BN - 1) / TC_BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (P.K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC_STAGES; s++) {
      tc_mbar_init(&full[s], 1);
      tc_mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      tc_mbar_init(&tmem_full[s], 1);
      tc_mbar_init(&tmem_empty[s], TC_EPI_WARPS);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // TMEM allocation is warp-collective; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_ptr)),
                 "n"(2 * TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tile_m = t % tiles_m, tile_n = t / tiles_m;
        for (int kb = 0; kb < num_kb; kb++, it++) {
          const int s = it % TC_STAGES, use = it / TC_STAGES;
          if (use > 0) tc_mbar_wait(&empty[s], (use - 1) & 1);
          unsigned char *st = tiles + s * TC_STAGE_BYTES;
          tc_mbar_expect_tx(&full[s], TC_STAGE_BYTES);
          tc_tma_load_2d(st + 0 * TC_TILE_BYTES, &map_a_big, kb * TC_BK, tile_m * TC_BM, &full[s]);
          tc_tma_load_2d(st + 1 * TC_TILE_BYTES, &map_a_small, kb * TC_BK, tile_m * TC_BM, &full[s]);
          tc_tma_load_2d(st + 2 * TC_TILE_BYTES, &map_b_big, kb * TC_BK, tile_n * TC_BN, &full[s]);
          tc_tma_load_2d(st + 3 * TC_TILE_BYTES, &map_b_small, kb * TC_BK, tile_n * TC_BN, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer
      const uint32_t idesc = tc_idesc();
      int it = 0, i = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, i++) {
        const int as = i & 1, ause = i >> 1;
        if (ause > 0) {  // the epilogue must have drained this accumulator
          tc_mbar_wait(&tmem_empty[as], (ause - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tmem_c = tmem_base + (uint32_t)(as * TC_TMEM_COLS);
        for (int kb = 0; kb < num_kb; kb++, it++) {
          const int s = it % TC_STAGES, use = it / TC_STAGES;
          tc_mbar_wait(&full[s], use & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = tc_smem_u32(tiles + s * TC_STAGE_BYTES);
          const uint64_t a_big = tc_make_desc(base), a_small = tc_make_desc(base + TC_TILE_BYTES);
          const uint64_t b_big = tc_make_desc(base + 2 * TC_TILE_BYTES), b_small = tc_make_desc(base + 3 * TC_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; k++) {  // UMMA_K = 8 tf32 = 32 bytes: advance the start address by 2 (x16 B)
            const uint64_t adv = (uint64_t)(k * 2);
            tc_mma(tmem_c, a_small + adv, b_big + adv, idesc, (kb | k) ? 1u : 0u);
            tc_mma(tmem_c, a_big + adv, b_small + adv, idesc, 1u);
            tc_mma(tmem_c, a_big + adv, b_big + adv, idesc, 1u);
          }
          tc_commit(&empty[s]);  // the slot is free once these MMAs have read it
        }
        tc_commit(&tmem_full[as]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue warps 2..9: TMEM lanes (warp % 4) * 32 .. + 31 (the hardware ties a warp to that lane quarter);
    // two warps share a quarter and take 64 of the tile's 128 columns each.  Everything arithmetic happens in the
    // accumulator's own layout (lane = row, 16 columns per tcgen05.ld): |q|^2 is lane-local, the 16 |y|^2 come as four
    // broadcast LDS.128 from a per-warp strip, and the minimum over the row's 32-column chunk is a lane-local running
    // min — ~4 instructions per element.  The tile then goes through a padded shared-memory transpose with 128-bit
    // accesses both ways so that global stores are whole 64-byte row segments (profiles/r02f: the previous epilogue
    // — scalar transpose, a shuffle, a REDUX and 64-bit address arithmetic per element, ~34 instructions per element on
    // two warps per scheduler — took ~10 k cycles per tile against ~4 k for the tile's MMAs).
    const int quarter = warp & 3;
    const int col_half = (warp - 2) >> 2;  // 0: columns 0..63, 1: columns 64..127
    float *tr = tr_all + (warp - 2) * TC_TR_FLOATS;  // [32 rows][TC_TR_PITCH] transpose strip, then 64 floats of |y|^2
    float *bn_w = tr + 32 * TC_TR_PITCH;
    const float INF = __int_as_float(0x7f800000);
    int i = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, i++) {
      const int tile_m = t % tiles_m, tile_n = t / tiles_m;
      const int as = i & 1, ause = i >> 1;
      const int row0 = tile_m * TC_BM + quarter * 32;  // this warp's 32 rows; lane l holds row row0 + l
      const int colw = tile_n * TC_BN + col_half * (TC_BN / 2);  // this warp's 64 columns
      const float an_mine = (P.l2 && P.a_norm && row0 + lane < P.M) ? P.a_norm[row0 + lane] : 0.f;
      if (P.l2) {  // |y|^2 of the warp's columns; +inf beyond N keeps those columns out of the chunk minima
        __syncwarp();
        bn_w[lane] = (P.b_norm && colw + lane < P.N) ? __ldg(P.b_norm + colw + lane) : (colw + lane < P.N ? 0.f : INF);
        bn_w[32 + lane] =
            (P.b_norm && colw + 32 + lane < P.N) ? __ldg(P.b_norm + colw + 32 + lane) : (colw + 32 + lane < P.N ? 0.f : INF);
        __syncwarp();
      }
      tc_mbar_wait(&tmem_full[as], ause & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float cmin_run = INF;
#pragma unroll 1
      for (int c0 = 0; c0 < TC_BN / 2; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr =
            tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TC_TMEM_COLS + col_half * (TC_BN / 2) + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c0 + 16 >= TC_BN / 2) {  // this warp's share of the accumulator is in registers: hand it back
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (lane == 0) tc_mbar_arrive(&tmem_empty[as]);
        }
        float r[16];
        if (P.l2) {
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            const float4 bn = *reinterpret_cast<const float4 *>(bn_w + c0 + j4 * 4);  // broadcast
            const float b[4] = {bn.x, bn.y, bn.z, bn.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
              // (|q|^2 + |y|^2) - 2 <q, y>, one rounding each, as faiss computes it; negative -> 0, NaN stays NaN
              float x = fmaf(-2.f, __uint_as_float(v[j4 * 4 + j]), an_mine + b[j]);
              asm("max.NaN.f32 %0, %0, 0f00000000;" : "+f"(x));
              r[j4 * 4 + j] = x;
              cmin_run = fminf(cmin_run, x);  // drops NaN
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; j++) r[j] = __uint_as_float(v[j]);
        }
#pragma unroll
        for (int j4 = 0; j4 < 4; j4++)
          *reinterpret_cast<float4 *>(tr + lane * TC_TR_PITCH + j4 * 4) = make_float4(r[j4 * 4], r[j4 * 4 + 1], r[j4 * 4 + 2], r[j4 * 4 + 3]);
        __syncwarp();
        // 4 lanes per row, 8 rows per instruction: every store is a 64-byte row segment
        const int pr = lane >> 2, pc = (lane & 3) * 4;
        const int col = colw + c0 + pc;
        float *dst = P.out + (size_t)(row0 + pr) * P.ldo + col;
        const size_t step = (size_t)8 * P.ldo;
        const bool col_ok = col < P.N;  // N % 4 == 0: a 4-column piece is inside or outside as a whole
#pragma unroll
        for (int tq = 0; tq < 4; tq++) {
          const float4 o = *reinterpret_cast<const float4 *>(tr + (tq * 8 + pr) * TC_TR_PITCH + pc);
          if (col_ok && row0 + tq * 8 + pr < P.M) *reinterpret_cast<float4 *>(dst) = o;
          dst += step;
        }
        __syncwarp();
        if (P.cmin && (c0 & 16)) {  // a 32-column chunk is complete
          if (row0 + lane < P.M) P.cmin[(size_t)(row0 + lane) * P.cmin_pitch + ((colw + c0 - 16) >> 5)] = cmin_run;
          cmin_run = INF;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * TC_TMEM_COLS)
                 : "memory");
  }
}

// small = x - tf32_truncate(x)  (exact in fp32); the "big" operand is x itself: the tensor core drops the
// low 13 mantissa bits on its own.
__global__ void tf32_residual_kernel(const float *__restrict__ x, float *__restrict__ small, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = x[i];
    float big = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    small[i] = v - big;
  }
}

cudaError_t launch_tf32_residual(const float *x, float *small, size_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  tf32_residual_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, small, n);
  return cudaGetLastError();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows][K] fp32, box = 32 floats x 128 rows, SWIZZLE_128B, out-of-bounds reads give zeros
static bool make_map(CUtensorMap *m, const float *ptr, int rows, int K) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// out[M][ldo] (columns 0..N-1) = L2^2 or inner product of rows of a (M x K) against rows of b (N x K).
// K % 4 == 0 and 16-byte aligned bases are required by the tensor maps.
cudaError_t launch_tc_gemm(const float *a, const float *a_small, const float *a_norm, const float *b,
                           const float *b_small, const float *b_norm, int M, int N, int K, float *out, int ldo, int l2,
                           float *cmin, int cmin_pitch, cudaStream_t st) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if ((K & 3) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15) || ((uintptr_t)a_small & 15) || ((uintptr_t)b_small & 15))
    return cudaErrorInvalidValue;
  // the epilogue stores 16-byte pieces: rows of out start on 16-byte boundaries and hold N rounded up to 4 columns
  if ((ldo & 3) || ((uintptr_t)out & 15) || ldo < ((N + 3) & ~3)) return cudaErrorInvalidValue;
  CUtensorMap ma, mas, mb, mbs;
  if (!make_map(&ma, a, M, K) || !make_map(&mas, a_small, M, K) || !make_map(&mb, b, N, K) ||
      !make_map(&mbs, b_small, N, K))
    return cudaErrorNotSupported;
  const size_t smem = (size_t)TC_STAGES * TC_STAGE_BYTES + (2 * TC_STAGES + 4) * sizeof(uint64_t) + 16 +
                      TC_EPI_WARPS * TC_TR_FLOATS * sizeof(float) + 1024;
  {
    cudaError_t e = ensure_dynamic_smem(tc_gemm_tf32x3_kernel, smem);
    if (e != cudaSuccess) return e;
  }
  if (cmin && (!l2 || cmin_pitch < ((N + TC_BN - 1) / TC_BN) * (TC_BN / 32))) return cudaErrorInvalidValue;
  TcGemmParams P;
  P.a_norm = a_norm;
  P.b_norm = b_norm;
  P.out = out;
  P.M = M;
  P.N = N;
  P.K = K;
  P.ldo = ldo;
  P.l2 = l2;
  P.cmin = cmin;
  P.cmin_pitch = cmin_pitch;
  const int num_tiles = ((N + TC_BN - 1) / TC_BN) * ((M + TC_BM - 1) / TC_BM);
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = num_tiles < sms ? num_tiles : sms;  // persistent: one CTA per SM
  tc_gemm_tf32x3_kernel<<<grid, TC_THREADS, smem, st>>>(ma, mas, mb, mbs, P);
  return cudaGetLastError();
}

}  // namespace gb
