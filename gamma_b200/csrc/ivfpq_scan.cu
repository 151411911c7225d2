// K2 — fused IVFPQ ADC scan: per-query lookup table in shared memory, validity filter
// inside the scan, streaming top-recall_num selection.  Replaces (reference file:line)
//   QueryTables::init_query / precompute_list_tables   index/impl/gamma_index_ivfpq.h:148-299
//   GammaIVFPQScanner::scan_list_with_table            index/impl/gamma_index_ivfpq.h:576-601
//   KnnSearchResults::add + heap_replace_top           index/impl/gamma_index_ivfpq.h:351-370
//   scan_one_list / the per-query probe loop           index/impl/gamma_index_ivfpq.cc:597-640, 790-818
//   RTInvertIndex::GetIvtList (list walk)              realtime/realtime_invert_index.cc:77-81
//
// Arithmetic restated for the GPU (identical algebra, fp32):
//   reference L2 :  dis = coarse_dis + SUM_m ( T[key][m][c_m] - 2 * (q_m . cb[m][c_m]) ),
//                   T[key][m][c] = |cb[m][c]|^2 + 2 * (centroid[key]_m . cb[m][c])     (faiss IndexIVFPQ.cpp:411-453)
//   here         :  dis = coarse_dis + t(p) + SUM_m lut[m][c_m],   lut[m][c] = -2 * (q_m . cb[m][c]),
//                   t(p) = SUM_m T[key][m][c_m(p)]  stored per posting at append time (4 B),
//   so the table depends on the QUERY only (one 32/64 KB table per query instead of one per
//   (query,list) pair) and is never written to HBM.
//   InnerProduct :  dis = q . centroid[key] + SUM_m (q_m . cb[m][c_m])                   (gamma_index_ivfpq.h:216-230)
//
// Shared-memory lookups, not HBM, bound this kernel (M lookups per posting).  The M = 32
// specialisation makes every warp-wide lookup bank-conflict free: the table is stored
// code-major [256][64] (entries for m = 0..31 duplicated at 32..63), lane l processes
// posting l of a 32-posting block and at step s reads sub-quantiser (l + s) mod 32 — word
// (l + s) of row `code`, i.e. bank (l + s) mod 32, distinct for the 32 lanes.  The device
// posting mirror stores each posting's code bytes pre-rotated by its lane so the byte for
// step s sits at a compile-time register position; one PRMT builds (code << 8 | lane*4) and
// the LDS carries 4*s as an immediate:  PRMT + LDS + FADD per lookup.
#include "common.cuh"
#include "kernels.h"

// dynamic shared memory at file scope so PTX can name it: its shared-window address is a link-time
// constant that ptxas folds into the LDS immediate (no per-lookup base add).
extern __shared__ __align__(16) unsigned char gb_scan_smem[];

namespace gb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_U = 2;  // 32-posting blocks per warp per round
constexpr int SCAN_ROUND_POSTINGS = SCAN_WARPS * SCAN_U * 32;

struct ProbeInfo {
  long long off;  // first posting of the list in the pools
  int len;        // postings visible to the scan (retrieve_idx_pos_)
  int rank;       // probe rank in the query's coarse ordering (tie-break order)
  float dis0;
};

// shared-memory carve-up (host mirrors this in scan_smem_bytes)
//  [lut][buf u64 cap][qs d floats][probe infos][blk prefix][misc]
__host__ __device__ inline size_t scan_lut_bytes(int M, int mode) {
  return mode == 1 ? (size_t)256 * 64 * 4 : (size_t)M * 257 * 4;
}

template <bool IP>
__device__ __forceinline__ float adc_generic(const float *lut, const uint8_t *codes, long long blk_base,
                                             int lane, int M, int chunk) {
  // layout 0: block of 32 postings, chunk-major: byte b of posting `lane` lives at
  //   blk_base + ((b / chunk) * 32 + lane) * chunk + (b % chunk)
  float acc = 0.f;
  int nwords = M >> 2;
  for (int w = 0; w < nwords; w++) {
    int b = w << 2;
    long long a = blk_base + ((long long)(b / chunk) * 32 + lane) * chunk + (b % chunk);
    uint32_t word = __ldg((const uint32_t *)(codes + a));
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int c = (word >> (8 * j)) & 0xff;
      acc += lut[(b + j) * 257 + c];
    }
  }
  return acc;
}

template <int IMM>
__device__ __forceinline__ float lds_lut(uint32_t off) {
  float v;
  // address = gb_scan_smem (constant) + off + IMM ; the table is the first thing in dynamic smem
  asm("{\n\t.reg .u32 a;\n\tmov.u32 a, gb_scan_smem;\n\tadd.u32 a, a, %1;\n\tld.shared.f32 %0, [a+%2];\n\t}"
      : "=f"(v)
      : "r"(off), "n"(IMM));
  return v;
}

// M = 32 conflict-free ADC (see header comment).  lut_lane = shared byte address of the table
// + lane*4 folded into the PRMT operand; w[0..7] = the posting's 32 pre-rotated code bytes.
template <int S>
__device__ __forceinline__ void adc_m32_word(uint32_t word, uint32_t lane4, float &a0, float &a1, float &a2, float &a3) {
  // (code_byte << 8) | lane4 : selector nibbles [3]=5 (zero) [2]=5 (zero) [1]=byte j [0]=4 (lane4)
  a0 += lds_lut<4 * (S + 0)>(__byte_perm(word, lane4, 0x5504));
  a1 += lds_lut<4 * (S + 1)>(__byte_perm(word, lane4, 0x5514));
  a2 += lds_lut<4 * (S + 2)>(__byte_perm(word, lane4, 0x5524));
  a3 += lds_lut<4 * (S + 3)>(__byte_perm(word, lane4, 0x5534));
}
__device__ __forceinline__ float adc_m32(uint32_t lane4, const uint32_t (&w)[8]) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  adc_m32_word<0>(w[0], lane4, a0, a1, a2, a3);
  adc_m32_word<4>(w[1], lane4, a0, a1, a2, a3);
  adc_m32_word<8>(w[2], lane4, a0, a1, a2, a3);
  adc_m32_word<12>(w[3], lane4, a0, a1, a2, a3);
  adc_m32_word<16>(w[4], lane4, a0, a1, a2, a3);
  adc_m32_word<20>(w[5], lane4, a0, a1, a2, a3);
  adc_m32_word<24>(w[6], lane4, a0, a1, a2, a3);
  adc_m32_word<28>(w[7], lane4, a0, a1, a2, a3);
  return (a0 + a1) + (a2 + a3);
}

template <bool IP, int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) ivfpq_scan_kernel(ScanParams P) {
  unsigned char *smem = gb_scan_smem;
  const int q = blockIdx.y;
  const int split = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int M = P.M, d = P.d, dsub = P.dsub;

  // ---- carve shared memory
  float *lut = reinterpret_cast<float *>(smem);
  size_t o = scan_lut_bytes(M, MODE);
  u64 *buf = reinterpret_cast<u64 *>(smem + o);
  o += (size_t)P.cap * sizeof(u64);
  float *qs = reinterpret_cast<float *>(smem + o);
  o += (size_t)((d + 3) & ~3) * sizeof(float);
  const int np_s = (P.nprobe - split + P.S - 1) / P.S;  // my probes: split, split+S, ...
  ProbeInfo *pinfo = reinterpret_cast<ProbeInfo *>(smem + ((o + 15) & ~(size_t)15));
  o = ((o + 15) & ~(size_t)15) + (size_t)P.max_np_s * sizeof(ProbeInfo);
  int *blk_prefix = reinterpret_cast<int *>(smem + o);
  o += (size_t)(P.max_np_s + 1) * sizeof(int);
  int *misc = reinterpret_cast<int *>(smem + ((o + 7) & ~(size_t)7));
  // misc: [0..1] tau (u64), [2] cnt, [4..67] warp_part
  BlockTopR topr;
  topr.buf = buf;
  topr.tau = reinterpret_cast<u64 *>(misc);
  topr.cnt = misc + 2;
  topr.warp_part = misc + 4;
  topr.cap = P.cap;
  topr.R = P.R;

  // ---- query to shared
  const float *xq = P.xq + (size_t)q * d;
  for (int i = tid; i < d; i += SCAN_THREADS) qs[i] = xq[i];
  if (tid == 0) {
    *topr.cnt = 0;
    *topr.tau = GB_KEY_MAX;
  }
  __syncthreads();

  // ---- per-query lookup table: q_m . cb[m][c]   (x -2 for L2)
  // pq_t is code-major [256][M][dsub]: lanes over m read consecutive dsub-float groups.
  {
    const float scale = IP ? 1.f : -2.f;
    const int total = 256 * M;
    for (int e = tid; e < total; e += SCAN_THREADS) {
      int c = e / M, m = e - c * M;
      const float *cb = P.pq_t + (size_t)e * dsub;
      const float *qm = qs + m * dsub;
      float ip = 0.f;
      if ((dsub & 3) == 0) {
        for (int j = 0; j < dsub; j += 4) {
          float4 v = __ldg(reinterpret_cast<const float4 *>(cb + j));
          ip = fmaf(qm[j], v.x, ip);
          ip = fmaf(qm[j + 1], v.y, ip);
          ip = fmaf(qm[j + 2], v.z, ip);
          ip = fmaf(qm[j + 3], v.w, ip);
        }
      } else {
        for (int j = 0; j < dsub; j++) ip = fmaf(qm[j], __ldg(cb + j), ip);
      }
      float v = scale * ip;
      if (MODE == 1) {
        lut[c * 64 + m] = v;
        lut[c * 64 + 32 + m] = v;
      } else {
        lut[m * 257 + c] = v;
      }
    }
  }

  // ---- my probes: list extents, dis0, block prefix
  long long my_postings = 0;
  for (int j = tid; j < np_s; j += SCAN_THREADS) {
    int p = split + j * P.S;
    int key = P.keys[(size_t)q * P.nprobe + p];
    ProbeInfo pi;
    pi.rank = p;
    pi.off = 0;
    pi.len = 0;
    pi.dis0 = 0.f;
    if (key >= 0 && key < P.nlist) {  // scan_one_list: key < 0 or >= nlist => skip (gamma_index_ivfpq.cc:602-609)
      pi.off = P.list_off[key];
      pi.len = P.list_len[key];
      if (IP) {
        const float *cen = P.centroids + (size_t)key * d;
        float s = 0.f;
        for (int i = 0; i < d; i++) s = fmaf(qs[i], __ldg(cen + i), s);
        pi.dis0 = s;
      } else {
        pi.dis0 = P.coarse_dis[(size_t)q * P.nprobe + p];
      }
    }
    pinfo[j] = pi;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int j = 0; j < np_s; j++) {
      blk_prefix[j] = acc;
      acc += (pinfo[j].len + 31) >> 5;
      my_postings += pinfo[j].len;
    }
    blk_prefix[np_s] = acc;
    if (P.scanned) atomicAdd(P.scanned, (unsigned long long)my_postings);
  }
  __syncthreads();

  const int total_blocks = blk_prefix[np_s];
  const int rounds = (total_blocks + SCAN_WARPS * SCAN_U - 1) / (SCAN_WARPS * SCAN_U);
  const uint32_t lane4 = lane * 4;
  const int prune_limit = P.cap - SCAN_ROUND_POSTINGS;
  int cur = 0;  // probe cursor (monotone per warp)

  for (int r = 0; r < rounds; r++) {
#pragma unroll
    for (int u = 0; u < SCAN_U; u++) {
      int g = (r * SCAN_U + u) * SCAN_WARPS + warp;
      if (g < total_blocks) {  // warp-uniform
        while (g >= blk_prefix[cur + 1]) cur++;
        const ProbeInfo pi = pinfo[cur];
        const int b = g - blk_prefix[cur];
        const int pos = b * 32 + lane;
        const long long pidx = pi.off + pos;
        bool ok = pos < pi.len;
        int id = -1;
        float dis = pi.dis0;
        if (MODE == 1) {
          // 32-posting block = 2 chunks of 512 B: chunk j of posting `lane` at (j*32 + lane)*16
          const uint8_t *blk = P.codes + (size_t)(pi.off + (long long)b * 32) * 32;
          uint4 c0 = ldg_nc_v4(blk + lane * 16);
          uint4 c1 = ldg_nc_v4(blk + 512 + lane * 16);
          if (ok) id = ldg_nc_s32(P.ids + pidx);
          if (!IP) dis += ok ? ldg_nc_f32(P.norms + pidx) : 0.f;
          ok = ok && id >= 0;
          if (ok && P.valid) ok = bitmap_test(P.valid, id);
          uint32_t w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          dis += adc_m32(lane4, w);
        } else {
          if (ok) id = ldg_nc_s32(P.ids + pidx);
          ok = ok && id >= 0;
          if (ok && P.valid) ok = bitmap_test(P.valid, id);
          if (ok) {
            if (!IP) dis += ldg_nc_f32(P.norms + pidx);
            long long blk_base = (pi.off + (long long)b * 32) * (long long)M;
            dis += adc_generic<IP>(lut, P.codes, blk_base, lane, M, P.chunk);
          }
        }
        uint32_t seq = ((uint32_t)pi.rank << GB_SEQ_POS_BITS) | (uint32_t)pos;
        u64 key = ((u64)dist_to_key32<IP>(dis) << 32) | seq;
        bool pass = ok && (dis == dis) && key < topr.threshold();
        topr.append_warp(pass, key);
      }
    }
    int over = *((volatile int *)topr.cnt) > prune_limit;
    if (__syncthreads_or(over)) topr.prune_collective();
  }
  topr.prune_collective();  // leaves min(cnt, R) survivors

  const int n_out = min(*((volatile int *)topr.cnt), P.R);
  u64 *out = P.cand + ((size_t)q * P.S + split) * P.R;
  for (int i = tid; i < P.R; i += SCAN_THREADS) out[i] = i < n_out ? buf[i] : GB_KEY_MAX;
}

size_t scan_smem_bytes(const ScanParams &P, int mode) {
  size_t o = scan_lut_bytes(P.M, mode);
  o += (size_t)P.cap * sizeof(u64);
  o += (size_t)((P.d + 3) & ~3) * sizeof(float);
  o = ((o + 15) & ~(size_t)15) + (size_t)P.max_np_s * sizeof(ProbeInfo);
  o += (size_t)(P.max_np_s + 1) * sizeof(int);
  o = ((o + 7) & ~(size_t)7) + (4 + 64) * sizeof(int);
  return o;
}

int scan_buffer_cap(int R) {
  // room for R survivors + one full round of admissions
  int need = R + SCAN_ROUND_POSTINGS;
  int cap = 1024;
  while (cap < need) cap <<= 1;
  return cap;  // <= 16 * SCAN_THREADS = 4096 (BlockTopR::prune_collective register budget)
}

template <bool IP, int MODE>
static cudaError_t launch_one(const ScanParams &P, cudaStream_t st) {
  size_t smem = scan_smem_bytes(P, MODE);
  static size_t configured[2][2] = {{0, 0}, {0, 0}};
  if (smem > configured[IP][MODE]) {
    cudaError_t e = cudaFuncSetAttribute(ivfpq_scan_kernel<IP, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[IP][MODE] = smem;
  }
  dim3 grid(P.S, P.n);
  ivfpq_scan_kernel<IP, MODE><<<grid, SCAN_THREADS, smem, st>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_ivfpq_scan(const ScanParams &P, int mode, cudaStream_t st) {
  if (P.is_ip) return mode == 1 ? launch_one<true, 1>(P, st) : launch_one<true, 0>(P, st);
  return mode == 1 ? launch_one<false, 1>(P, st) : launch_one<false, 0>(P, st);
}

}  // namespace gb
