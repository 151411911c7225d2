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
#include "scan_common.cuh"

namespace gb {

// generic kernel: lock-step rounds of SCAN_U blocks per warp
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_U = 2;  // 32-posting blocks per warp per round
constexpr int SCAN_ROUND_POSTINGS = SCAN_WARPS * SCAN_U * 32;
// M = 32 kernel: 12 warps, 3 CTAs per SM (36 resident warps), sync point every M32_U blocks per warp
constexpr int M32_U = 8;

// shared-memory carve-up (host mirrors this in scan_smem_bytes)
//  [lut][buf u64 cap][qs d floats][probe infos][blk prefix][misc]
__host__ __device__ inline size_t scan_lut_bytes(int M, int mode) {
  return mode == 1 ? (size_t)256 * 64 * 4 : mode == 2 ? (size_t)256 * 96 * 4 : (size_t)M * 257 * 4;
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


// ---------------------------------------------------------------------------------------------
// shared-memory carve-up, identical for both kernels (host mirrors it in scan_smem_bytes):
//   [lut][buf u64 cap][qs d floats][probe infos][blk prefix][misc 68 ints][mbar 2 x u64]
// ---------------------------------------------------------------------------------------------
struct ScanSmem {
  float *lut;
  u64 *buf;
  float *qs;
  ProbeInfo *pinfo;
  int *blk_prefix;
  int *misc;
  unsigned long long *mbar;
};

// bytes of the [ProbeInfo x max_np_s][blk_prefix x (max_np_s + 1)] region, padded to 16 so that one bulk copy
// can fill it (the v2 kernel receives it precomputed from probe_setup_kernel)
__host__ __device__ inline size_t scan_probe_bytes(int max_np_s) {
  size_t b = (size_t)max_np_s * sizeof(ProbeInfo) + (size_t)(max_np_s + 1) * sizeof(int);
  return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ ScanSmem carve(unsigned char *smem, const ScanParams &P, int mode) {
  ScanSmem S;
  S.lut = reinterpret_cast<float *>(smem);
  size_t o = scan_lut_bytes(P.M, mode);
  S.buf = reinterpret_cast<u64 *>(smem + o);
  o += (size_t)P.cap * sizeof(u64);
  S.qs = reinterpret_cast<float *>(smem + o);
  o += (size_t)((P.d + 3) & ~3) * sizeof(float);
  o = (o + 15) & ~(size_t)15;
  S.pinfo = reinterpret_cast<ProbeInfo *>(smem + o);
  S.blk_prefix = reinterpret_cast<int *>(smem + o + (size_t)P.max_np_s * sizeof(ProbeInfo));
  o += scan_probe_bytes(P.max_np_s);
  S.misc = reinterpret_cast<int *>(smem + o);  // [0..1] tau (u64), [2] cnt, [4..67] warp_part, [68..70] round flags
  o += (4 + 64 + 4) * sizeof(int);
  S.mbar = reinterpret_cast<unsigned long long *>(smem + o);
  return S;
}

size_t scan_probe_bytes_host(int max_np_s) { return scan_probe_bytes(max_np_s); }

size_t scan_smem_bytes(const ScanParams &P, int mode) {
  size_t o = scan_lut_bytes(P.M, mode);
  o += (size_t)P.cap * sizeof(u64);
  o += (size_t)((P.d + 3) & ~3) * sizeof(float);
  o = ((o + 15) & ~(size_t)15) + scan_probe_bytes(P.max_np_s);
  o += (4 + 64 + 4) * sizeof(int);
  o += 2 * sizeof(unsigned long long);
  return o;
}

// my probes (split, split+S, ...): list extents, dis0, prefix of 32-posting blocks.  collective.
template <bool IP>
__device__ __forceinline__ int setup_probes(const ScanParams &P, const ScanSmem &S, const float *q_glob, int q,
                                            int split) {
  const int tid = threadIdx.x;
  const int d = P.d;
  const int np_s = (P.nprobe - split + P.S - 1) / P.S;
  for (int j = tid; j < np_s; j += blockDim.x) {
    int p = split + j * P.S;
    int key = P.keys[(size_t)q * P.nprobe + p];
    ProbeInfo pi;
    pi.rank = p;
    pi.off = 0;
    pi.len = 0;
    pi.dis0 = 0.f;
    if (key >= 0 && key < P.nlist) {  // scan_one_list: key < 0 or >= nlist => skip (gamma_index_ivfpq.cc:602-609)
      load_list_extent(P.list_off, P.list_len, key, pi.off, pi.len);
      if (IP) {  // dis0 = <q, centroid>  (precompute_list_tables_IP, gamma_index_ivfpq.h:216-230)
        const float *cen = P.centroids + (size_t)key * d;
        float s = 0.f;
        for (int i = 0; i < d; i++) s = fmaf(__ldg(q_glob + i), __ldg(cen + i), s);
        pi.dis0 = s;
      } else {
        pi.dis0 = P.coarse_dis[(size_t)q * P.nprobe + p];
      }
    }
    S.pinfo[j] = pi;
  }
  __syncthreads();
  if (tid < 32) {  // exclusive prefix of the per-list block counts (warp scan, 32 lists per pass)
    int carry = 0;
    long long my_postings = 0;
    for (int j0 = 0; j0 < np_s; j0 += 32) {
      const int j = j0 + tid;
      const int len = j < np_s ? S.pinfo[j].len : 0;
      const int nb = (len + 31) >> 5;
      int incl = nb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(GB_FULL, incl, o);
        if (tid >= o) incl += v;
      }
      if (j < np_s) S.blk_prefix[j] = carry + incl - nb;
      carry += __shfl_sync(GB_FULL, incl, 31);
      my_postings += len;
    }
    my_postings = __reduce_add_sync(GB_FULL, (unsigned)my_postings);
    if (tid == 0) {
      S.blk_prefix[np_s] = carry;
      if (P.scanned) atomicAdd(P.scanned, (unsigned long long)my_postings);
    }
  }
  __syncthreads();
  return S.blk_prefix[np_s];
}

__device__ __forceinline__ BlockTopR make_topr(const ScanSmem &S, const ScanParams &P) {
  BlockTopR t;
  t.buf = S.buf;
  t.tau = reinterpret_cast<u64 *>(S.misc);
  t.cnt = S.misc + 2;
  t.warp_part = S.misc + 4;
  t.cap = P.cap;
  t.R = P.R;
  return t;
}

template <int PER>
__device__ __forceinline__ void write_survivors(BlockTopR &topr, const ScanParams &P, int q, int split) {
  topr.prune_collective<PER>();  // leaves min(cnt, R) survivors
  const int n_out = min(*((volatile int *)topr.cnt), P.R);
  u64 *out = P.cand + ((size_t)q * P.S + split) * P.R;
  for (int i = threadIdx.x; i < P.R; i += blockDim.x) out[i] = i < n_out ? topr.buf[i] : GB_KEY_MAX;
}

// =============================================================================================
// generic kernel: any M % 4 == 0, classic [M][257] table, lanes = postings (bank conflicts random)
// =============================================================================================
template <bool IP, int PER>
__global__ void __launch_bounds__(SCAN_THREADS) ivfpq_scan_generic_kernel(ScanParams P) {
  const int q = blockIdx.y, split = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = P.M, d = P.d, dsub = P.dsub;
  ScanSmem S = carve(gb_scan_smem, P, 0);
  BlockTopR topr = make_topr(S, P);
  const float *xq = P.xq + (size_t)q * d;
  for (int i = tid; i < d; i += SCAN_THREADS) S.qs[i] = xq[i];
  if (tid == 0) {
    *topr.cnt = 0;
    *topr.tau = GB_KEY_MAX;
  }
  __syncthreads();
  {  // per-query table  q_m . cb[m][c]  (x -2 for L2); pq_t is code-major [256][M][dsub]
    const float scale = IP ? 1.f : -2.f;
    const int total = 256 * M;
    for (int e = tid; e < total; e += SCAN_THREADS) {
      int c = e / M, m = e - c * M;
      const float *cb = P.pq_t + (size_t)e * dsub;
      const float *qm = S.qs + m * dsub;
      float ip = 0.f;
      for (int j = 0; j < dsub; j++) ip = fmaf(qm[j], __ldg(cb + j), ip);
      S.lut[m * 257 + c] = scale * ip;
    }
  }
  const int total_blocks = setup_probes<IP>(P, S, xq, q, split);
  const int rounds = (total_blocks + SCAN_WARPS * SCAN_U - 1) / (SCAN_WARPS * SCAN_U);
  const int prune_limit = P.cap - SCAN_ROUND_POSTINGS;
  int cur = 0;
  for (int r = 0; r < rounds; r++) {
#pragma unroll
    for (int u = 0; u < SCAN_U; u++) {
      int g = (r * SCAN_U + u) * SCAN_WARPS + warp;
      if (g < total_blocks) {  // warp-uniform
        while (g >= S.blk_prefix[cur + 1]) cur++;
        const ProbeInfo pi = S.pinfo[cur];
        const int b = g - S.blk_prefix[cur];
        const int pos = b * 32 + lane;
        const long long pidx = pi.off + pos;
        bool ok = pos < pi.len;
        int id = -1;
        float dis = pi.dis0;
        if (ok) id = ldg_nc_s32(P.ids + pidx);
        ok = ok && id >= 0;
        if (ok && P.valid) ok = (long long)id < P.valid_bits && bitmap_test(P.valid, id);
        if (ok) {
          if (!IP) dis += ldg_nc_f32(P.norms + pidx);
          long long blk_base = (pi.off + (long long)b * 32) * (long long)M;
          dis += adc_generic<IP>(S.lut, P.codes, blk_base, lane, M, P.chunk);
        }
        uint32_t seq = ((uint32_t)pi.rank << GB_SEQ_POS_BITS) | (uint32_t)pos;
        u64 key = ((u64)dist_to_key32<IP>(dis) << 32) | seq;
        bool pass = ok && (dis == dis) && key < topr.threshold();
        topr.append_warp(pass, key);
      }
    }
    int over = *((volatile int *)topr.cnt) > prune_limit;
    if (__syncthreads_or(over)) topr.prune_collective<PER>();
  }
  write_survivors<PER>(topr, P, q, split);
}

// K2b — probe tables for the M = 64 kernel: one warp per (query, split) writes, in the kernel's shared-memory
// layout, [ProbeInfo x max_np_s][exclusive prefix of 32-posting block counts x (max_np_s + 1)]:
// scan_one_list's list lookup (gamma_index_ivfpq.cc:597-640) and dis0 of precompute_list_tables
// (gamma_index_ivfpq.h:216-230, 236-299) hoisted out of the scan.
__global__ void __launch_bounds__(256) probe_setup_kernel(ScanParams P) {
  const int item = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (item >= (P.n_items > 0 ? P.n_items : P.n * P.S)) return;
  int q, split, nsp;
  if (P.n_items > 0) {
    if (item < P.n_full) {
      q = item, split = 0, nsp = 1;
    } else {
      const int t = item - P.n_full;
      q = P.n_full + t / P.s_tail, split = t % P.s_tail, nsp = P.s_tail;
    }
  } else {
    q = item / P.S, split = item - q * P.S, nsp = P.S;
  }
  const int np_s = (P.nprobe - split + nsp - 1) / nsp;
  const size_t pbytes = scan_probe_bytes(P.max_np_s);
  unsigned char *dst = P.probe_g + (size_t)item * pbytes;
  ProbeInfo *pinfo = reinterpret_cast<ProbeInfo *>(dst);
  int *prefix = reinterpret_cast<int *>(dst + (size_t)P.max_np_s * sizeof(ProbeInfo));
  const float *xq = P.xq + (size_t)q * P.d;
  int carry = 0;
  unsigned my_postings = 0;
  for (int j0 = 0; j0 < np_s; j0 += 32) {
    const int j = j0 + lane;
    ProbeInfo pi;
    pi.off = 0, pi.len = 0, pi.rank = 0, pi.dis0 = 0.f;
    if (j < np_s) {
      const int p = split + j * nsp;
      const int key = P.keys[(size_t)q * P.nprobe + p];
      pi.rank = p;
      if (key >= 0 && key < P.nlist) {  // scan_one_list: key < 0 or >= nlist => skip (gamma_index_ivfpq.cc:602-609)
        load_list_extent(P.list_off, P.list_len, key, pi.off, pi.len);
        if (P.is_ip) {  // dis0 = <q, centroid>
          const float *cen = P.centroids + (size_t)key * P.d;
          float s = 0.f;
          for (int i = 0; i < P.d; i++) s = fmaf(__ldg(xq + i), __ldg(cen + i), s);
          pi.dis0 = s;
        } else {
          pi.dis0 = P.coarse_dis[(size_t)q * P.nprobe + p];
        }
      }
      pinfo[j] = pi;
    }
    const int nb = (pi.len + 31) >> 5;
    int incl = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(GB_FULL, incl, o);
      if (lane >= o) incl += v;
    }
    if (j < np_s) prefix[j] = carry + incl - nb;
    carry += __shfl_sync(GB_FULL, incl, 31);
    my_postings += (unsigned)pi.len;
  }
  my_postings = __reduce_add_sync(GB_FULL, my_postings);
  if (lane == 0) {
    prefix[np_s] = carry;
    if (P.scanned) atomicAdd(P.scanned, (unsigned long long)my_postings);
  }
}

cudaError_t launch_probe_setup(const ScanParams &P, cudaStream_t st) {
  const int items = P.n_items > 0 ? P.n_items : P.n * P.S;
  probe_setup_kernel<<<(items + 7) / 8, 256, 0, st>>>(P);
  return cudaGetLastError();
}

// K2a — per-query lookup tables for the M = 32 kernel: lut_g[q][c][m] = lut_g[q][c][32 + m] =
// scale * <q_m, cb[m][c]>  (scale = -2 for L2, +1 for InnerProduct): QueryTables::init_query /
// ProductQuantizer::compute_inner_prod_table (gamma_index_ivfpq.h:148-168, faiss ProductQuantizer.cpp:504-530).
__global__ void __launch_bounds__(256) lut_build_m32_kernel(const float *__restrict__ xq, const float *__restrict__ pq_t,
                                                            float *__restrict__ lut_g, int d, int dsub, float scale) {
  __shared__ float qs[1024];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < d; i += 256) qs[(i % dsub) * 32 + (i / dsub)] = xq[(size_t)q * d + i];
  __syncthreads();
  float *out = lut_g + (size_t)q * 16384;
  if (dsub == 4) {
    float q0 = qs[lane], q1 = qs[32 + lane], q2 = qs[64 + lane], q3 = qs[96 + lane];
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
      const int e = tid + i * 256;
      const float4 v = __ldg(reinterpret_cast<const float4 *>(pq_t) + e);
      float ip = q0 * v.x;
      ip = fmaf(q1, v.y, ip);
      ip = fmaf(q2, v.z, ip);
      ip = fmaf(q3, v.w, ip);
      const int c = e >> 5;
      out[c * 64 + lane] = scale * ip;
      out[c * 64 + 32 + lane] = scale * ip;
    }
  } else {
    for (int e = tid; e < 8192; e += 256) {
      const float *cb = pq_t + (size_t)e * dsub;
      float ip = 0.f;
      for (int j = 0; j < dsub; j++) ip = fmaf(qs[j * 32 + lane], __ldg(cb + j), ip);
      const int c = e >> 5;
      out[c * 64 + lane] = scale * ip;
      out[c * 64 + 32 + lane] = scale * ip;
    }
  }
}

cudaError_t launch_lut_build_m32(const float *xq, const float *pq_t, float *lut_g, int n, int d, int dsub, int is_ip,
                                 cudaStream_t st) {
  if (d > 1024) return cudaErrorInvalidValue;
  lut_build_m32_kernel<<<n, 256, 0, st>>>(xq, pq_t, lut_g, d, dsub, is_ip ? 1.f : -2.f);
  return cudaGetLastError();
}

int scan_buffer_cap(int R) {
  // room for R survivors + one full round of admissions (generic kernel's lock-step rounds)
  int need = R + SCAN_ROUND_POSTINGS;
  int cap = 1024;
  while (cap < need) cap <<= 1;
  return cap;  // <= 16 * SCAN_THREADS = 4096 (BlockTopR::prune_collective register budget)
}

template <typename K>
static cudaError_t launch_kernel(K kernel, const ScanParams &P, int mode, int threads, size_t *configured,
                                 cudaStream_t st) {
  size_t smem = scan_smem_bytes(P, mode);
  (void)configured;
  cudaError_t e = ensure_dynamic_smem(kernel, smem);
  if (e != cudaSuccess) return e;
  dim3 grid(P.S, P.n);
  if (P.n_items > 0) grid = dim3(P.n_items, 1);
  kernel<<<grid, threads, smem, st>>>(P);
  return cudaGetLastError();
}

// =============================================================================================
// M = 64 kernel (the reference's default nsubvector; BASELINE config C3 = PQ64x8): one CTA per (query, split), probe
// table + lookup table by TMA bulk copies on one mbarrier, static per-warp split, next block loaded into the code
// registers the address formation just freed, L2 prefetch pf blocks ahead.  Against the M = 32 kernel:
//  * table [256 codes][96 words] = 96 KB: words 64..95 duplicate 0..31, lane l reads word (l + s) of row `code` at step
//    s = 0..63 — bank (l + s) mod 32, conflict free; the mirror pre-rotates the 64 code bytes of posting i by i
//    (LAYOUT_M64_ROT) so that step s uses stored byte s;
//  * the row pitch (384 B) is not a power of two, so the address is PRMT (byte extract) + IMAD (byte * 384 + lane * 4)
//    + LDS + FADD = 4 instructions per lookup; IMAD runs on the FMA pipe next to PRMT on the ALU pipe;
//  * a block of 32 postings is 2 KB of codes (4 chunks of 16 B per posting): 16 code registers, refilled in two
//    halves as soon as the addresses of that half have been formed;
//  * 384 threads, 2 CTAs per SM (2 x 108 KB of shared memory), candidate buffer of 1024 keys.
// Validated on hardware in round 2 (tests/test_ivfpq_gpu.py::test_m64_default_kernel_parity); 0.59 of the measured HBM
// peak on config C3.
// =============================================================================================
template <bool IP, bool HAS_VALID, int WARPS, int PER>
__device__ __forceinline__ void scan_loop_m64(const ScanParams &P, const ScanSmem &S, BlockTopR &topr,
                                              const int total_blocks, const int np_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lane4 = lane * 4;
  const int per_warp = (total_blocks + WARPS - 1) / WARPS;
  const int w0 = min(total_blocks, warp * per_warp);
  int left = min(total_blocks, w0 + per_warp) - w0;  // blocks this warp still has to LOAD
  const int soft_limit = P.cap - WARPS * 32;
  volatile int *flags = S.misc + 68;
  const int pf = P.pf_blocks;
  // ids at or beyond the bitmap's size (appended after this search began) and negative ids (dead / padding) fail
  const uint32_t valid_lim = (uint32_t)(P.valid_bits < 0x7fffffffLL ? P.valid_bits : 0x7fffffffLL);

  int pj = 0;
  if (left > 0)
    while (S.blk_prefix[pj + 1] <= w0) pj++;
  int bl = 0, len = 0;
  uint32_t seq0 = 0;
  float dis0 = 0.f;
  const uint8_t *cptr = nullptr, *hptr = nullptr;  // this lane's 16 B of chunk 0 of the next block / chunk 2 of the block being loaded
  const int *iptr = nullptr;
  const float *nptr = nullptr;
  // L2 prefetch streams, one 128 B line per lane and block: lanes 0..15 the 2 KB of codes, lane 16 ids, lane 17 t(p)
  const char *pfp = nullptr;
  const uint32_t pf_stride = lane < 16 ? 2048u : 128u;
  const bool pf_lane = pf > 0 && lane < (IP ? 17 : 18);
  auto stream_base = [&](const ProbeInfo &pi, int b_start) -> const char * {
    const long long first = pi.off + (long long)b_start * 32;
    return lane < 16 ? reinterpret_cast<const char *>(P.codes + (size_t)first * 64) + lane * 128
                     : lane == 16 ? reinterpret_cast<const char *>(P.ids + first)
                                  : reinterpret_cast<const char *>(P.norms + first);
  };
  auto open_list = [&](int j, int b_start) {
    const ProbeInfo pi = S.pinfo[j];
    bl = ((pi.len + 31) >> 5) - b_start;
    dis0 = pi.dis0;
    seq0 = ((uint32_t)pi.rank << GB_SEQ_POS_BITS) + (uint32_t)(b_start * 32 + lane);
    len = pi.len - (b_start * 32 + lane);
    const long long first = pi.off + (long long)b_start * 32;
    cptr = P.codes + (size_t)first * 64 + lane * 16;
    iptr = P.ids + first + lane;
    nptr = P.norms + first + lane;
    if (pf_lane) {
      pfp = stream_base(pi, b_start) + (size_t)pf * pf_stride;
      if (left > bl && j + 1 < np_s) {  // head of the next list this warp will walk
        const ProbeInfo pn = S.pinfo[j + 1];
        const int nb = min(min((pn.len + 31) >> 5, pf), left - bl);
        const char *h = stream_base(pn, 0);
#pragma unroll 1
        for (int b = 0; b < nb; b++) l2_prefetch_line(h + (size_t)b * pf_stride);
      }
    }
  };
  if (left > 0) {
    if (pf_lane) {
      const ProbeInfo pi = S.pinfo[pj];
      const int b0 = w0 - S.blk_prefix[pj];
      const int nb = min(min(((pi.len + 31) >> 5) - b0, pf), left);
      const char *h = stream_base(pi, b0);
#pragma unroll 1
      for (int b = 0; b < nb; b++) l2_prefetch_line(h + (size_t)b * pf_stride);
    }
    open_list(pj, w0 - S.blk_prefix[pj]);
  }

  // the block in flight: 64 pre-rotated code bytes (chunks 0,1 in c0..c7, chunks 2,3 in c8..c15), vid, t(p), dis0, seq
  uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
  uint32_t c8 = 0, c9 = 0, c10 = 0, c11 = 0, c12 = 0, c13 = 0, c14 = 0, c15 = 0;
  int id_n = -1;
  float nrm_n = 0.f, base_n = 0.f;
  uint32_t seq_n = 0xffffffffu;
  bool hi_pending = false;  // issue_lo loaded a block whose chunks 2,3 are still to be requested
  auto issue_lo = [&]() {
    hi_pending = false;
    if (left > 0) {  // warp-uniform
      while (bl == 0) open_list(++pj, 0);
      const uint4 v0 = ldg_nc_v4(cptr);
      const uint4 v1 = ldg_nc_v4(cptr + 512);
      c0 = v0.x, c1 = v0.y, c2 = v0.z, c3 = v0.w, c4 = v1.x, c5 = v1.y, c6 = v1.z, c7 = v1.w;
      hptr = cptr + 1024;
      hi_pending = true;
      seq_n = seq0;
      base_n = dis0;
      id_n = -1;
      nrm_n = 0.f;
      if (len > 0) {
        id_n = ldg_nc_s32(iptr);
        if (!IP) nrm_n = ldg_nc_f32(nptr);
      }
      if (pf_lane) {
        if (bl > pf) l2_prefetch_line(pfp);
        pfp += pf_stride;
      }
      cptr += 2048;
      iptr += 32;
      nptr += 32;
      seq0 += 32;
      len -= 32;
      bl--;
      left--;
    } else {
      seq_n = 0xffffffffu;
    }
  };
  auto issue_hi = [&]() {
    if (hi_pending) {  // warp-uniform
      const uint4 v2 = ldg_nc_v4(hptr);
      const uint4 v3 = ldg_nc_v4(hptr + 512);
      c8 = v2.x, c9 = v2.y, c10 = v2.z, c11 = v2.w, c12 = v3.x, c13 = v3.y, c14 = v3.z, c15 = v3.w;
    }
  };

  u64 skey = 0;
  bool spend = false;
  auto try_append = [&](bool pass, u64 key) -> bool {
    const unsigned m = __ballot_sync(GB_FULL, pass);
    if (m == 0) return false;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(topr.cnt, __popc(m));
    base = __shfl_sync(GB_FULL, base, leader);
    const int slot = base + __popc(m & ((1u << lane) - 1u));
    bool pending = pass;
    if (pass && slot < topr.cap) {
      topr.buf[slot] = key;
      pending = false;
    }
    spend = pending;
    skey = key;
    return __any_sync(GB_FULL, pending);
  };

  bool stalled = false;
  issue_lo();
  issue_hi();
  int round = 0;
  for (;;) {
    if (stalled) stalled = try_append(spend && skey < topr.threshold(), skey);  // after a prune
    bool over = false;
    while (!stalled && !over && seq_n != 0xffffffffu) {  // warp-uniform
      uint32_t a[16];
      float s0, s1, s2, s3;
      // address of table word (lane + step) of row `code`: code * 384 + lane * 4 (+ 0x400 + 4 * step in the LDS immediate)
#define GB64_ADDR4(W, I)                                   \
  a[I + 0] = prmt_v(W, 0u, 0x4440) * 384u + lane4;         \
  a[I + 1] = prmt_v(W, 0u, 0x4441) * 384u + lane4;         \
  a[I + 2] = prmt_v(W, 0u, 0x4442) * 384u + lane4;         \
  a[I + 3] = prmt_v(W, 0u, 0x4443) * 384u + lane4;
#define GB64_LOOK4(I, O)             \
  s0 += lds_raw<O + 0>(a[I + 0]);    \
  s1 += lds_raw<O + 1>(a[I + 1]);    \
  s2 += lds_raw<O + 2>(a[I + 2]);    \
  s3 += lds_raw<O + 3>(a[I + 3]);
      // quarter 0: steps 0..15 from chunk 0
      GB64_ADDR4(c0, 0) GB64_ADDR4(c1, 4) GB64_ADDR4(c2, 8) GB64_ADDR4(c3, 12)
      s0 = lds_raw<0>(a[0]), s1 = lds_raw<1>(a[1]), s2 = lds_raw<2>(a[2]), s3 = lds_raw<3>(a[3]);
      GB64_LOOK4(4, 4) GB64_LOOK4(8, 8) GB64_LOOK4(12, 12)
      // quarter 1: steps 16..31 from chunk 1; c0..c7 are dead afterwards -> first half of the next block
      GB64_ADDR4(c4, 0) GB64_ADDR4(c5, 4) GB64_ADDR4(c6, 8) GB64_ADDR4(c7, 12)
      const int id = id_n;
      const uint32_t seq = seq_n;
      const float nb = base_n + nrm_n;
      uint32_t vw = 0xffffffffu;
      if (HAS_VALID) vw = (uint32_t)id < valid_lim ? __ldg(P.valid + (id >> 5)) : 0u;
      issue_lo();
      GB64_LOOK4(0, 16) GB64_LOOK4(4, 20) GB64_LOOK4(8, 24) GB64_LOOK4(12, 28)
      // quarter 2: steps 32..47 from chunk 2
      GB64_ADDR4(c8, 0) GB64_ADDR4(c9, 4) GB64_ADDR4(c10, 8) GB64_ADDR4(c11, 12)
      GB64_LOOK4(0, 32) GB64_LOOK4(4, 36) GB64_LOOK4(8, 40) GB64_LOOK4(12, 44)
      // quarter 3: steps 48..63 from chunk 3; c8..c15 are dead afterwards -> second half of the next block
      GB64_ADDR4(c12, 0) GB64_ADDR4(c13, 4) GB64_ADDR4(c14, 8) GB64_ADDR4(c15, 12)
      issue_hi();
      GB64_LOOK4(0, 48) GB64_LOOK4(4, 52) GB64_LOOK4(8, 56) GB64_LOOK4(12, 60)
#undef GB64_ADDR4
#undef GB64_LOOK4
      const uint32_t tau_hi = *((volatile uint32_t *)topr.tau + 1);
      over = *((volatile int *)topr.cnt) > soft_limit;
      const float dis = nb + ((s0 + s1) + (s2 + s3));
      bool ok = id >= 0;
      if (HAS_VALID) ok = ok && ((vw >> (id & 31)) & 1u);
      const uint32_t k32 = dist_to_key32<IP>(dis);
      const bool pass = ok && (dis == dis) && k32 <= tau_hi;
      if (__any_sync(GB_FULL, pass)) {
        const u64 key = ((u64)k32 << 32) | seq;
        stalled = try_append(pass && key < topr.threshold(), key);
        over = over || *((volatile int *)topr.cnt) > soft_limit;
      }
    }
    const bool more = stalled || seq_n != 0xffffffffu;
    over = over || stalled;
    const int slot = round % 3;
    if (lane == 0 && (more || over)) atomicOr((int *)&flags[slot], (over ? 1 : 0) | (more ? 2 : 0));
    __syncthreads();
    const int v = flags[slot];
    if (threadIdx.x == 0) flags[(round + 2) % 3] = 0;
    round++;
    if (v & 1) topr.prune_collective<PER>();
    if (!(v & 2)) break;
  }
}

template <bool IP, int THREADS, int MINB, int PER>
__global__ void __launch_bounds__(THREADS, MINB) ivfpq_scan_m64_kernel(ScanParams P) {
  constexpr int WARPS = THREADS / 32;
  int q, split, nsp;
  size_t item;
  if (P.n_items > 0) {
    item = blockIdx.x;
    if ((int)blockIdx.x < P.n_full) {
      q = blockIdx.x, split = 0, nsp = 1;
    } else {
      const int t = blockIdx.x - P.n_full;
      q = P.n_full + t / P.s_tail, split = t % P.s_tail, nsp = P.s_tail;
    }
  } else {
    q = blockIdx.y;
    split = blockIdx.x, nsp = P.S;
    item = (size_t)q * P.S + split;
  }
  const int tid = threadIdx.x;
  ScanSmem S = carve(gb_scan_smem, P, 2);
  BlockTopR topr = make_topr(S, P);
  if (smem_u32(gb_scan_smem) != GB_SMEM_RESERVED) __trap();
  const int np_s = (P.nprobe - split + nsp - 1) / nsp;
  if (tid == 0) {
    *topr.cnt = 0;
    *topr.tau = GB_KEY_MAX;
    S.misc[68] = S.misc[69] = S.misc[70] = 0;
    mbar_init(&S.mbar[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t pbytes = (uint32_t)scan_probe_bytes(P.max_np_s);
    const char *src = reinterpret_cast<const char *>(P.lut_g) + (size_t)q * 98304;
    mbar_expect_tx(&S.mbar[0], 98304u + pbytes);
    tma_bulk_g2s(S.pinfo, P.probe_g + item * pbytes, pbytes, &S.mbar[0]);
#pragma unroll
    for (int i = 0; i < 6; i++)
      tma_bulk_g2s(reinterpret_cast<char *>(S.lut) + i * 16384, src + i * 16384, 16384u, &S.mbar[0]);
  }
  __syncthreads();
  mbar_wait(&S.mbar[0], 0);
  const int total_blocks = S.blk_prefix[np_s];
  if (P.valid) scan_loop_m64<IP, true, WARPS, PER>(P, S, topr, total_blocks, np_s);
  else scan_loop_m64<IP, false, WARPS, PER>(P, S, topr, total_blocks, np_s);
  write_survivors<PER>(topr, P, q, split);
  if (split == 0)
    for (int i = nsp * P.R + tid; i < P.S * P.R; i += THREADS) P.cand[(size_t)q * P.S * P.R + i] = GB_KEY_MAX;
}

// per-query tables for the M = 64 kernel: lut_g[q][c][w] = scale * <q_m, cb[m][c]>, m = w mod 64, w < 96
__global__ void __launch_bounds__(256) lut_build_m64_kernel(const float *__restrict__ xq, const float *__restrict__ pq_t,
                                                            float *__restrict__ lut_g, int d, int dsub, float scale) {
  __shared__ float qs[1024];
  const int q = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < d; i += 256) qs[i] = xq[(size_t)q * d + i];
  __syncthreads();
  float *out = lut_g + (size_t)q * (256 * 96);
  for (int e = tid; e < 256 * 64; e += 256) {  // e = c * 64 + m: pq_t is code-major [256][M][dsub]
    const int c = e >> 6, m = e & 63;
    const float *cb = pq_t + (size_t)e * dsub;
    const float *qm = qs + m * dsub;
    float ip = 0.f;
    for (int j = 0; j < dsub; j++) ip = fmaf(qm[j], __ldg(cb + j), ip);
    const float v = scale * ip;
    out[c * 96 + m] = v;
    if (m < 32) out[c * 96 + 64 + m] = v;
  }
}

cudaError_t launch_lut_build_m64(const float *xq, const float *pq_t, float *lut_g, int n, int d, int dsub, int is_ip,
                                 cudaStream_t st) {
  if (d > 1024) return cudaErrorInvalidValue;
  lut_build_m64_kernel<<<n, 256, 0, st>>>(xq, pq_t, lut_g, d, dsub, is_ip ? 1.f : -2.f);
  return cudaGetLastError();
}

template <int PER>
static cudaError_t launch_m64(const ScanParams &P, cudaStream_t st) {
  static size_t conf[2] = {0, 0};
  constexpr int MINB = PER == 4 ? 2 : 1;  // the 16-keys-per-thread select needs the registers
  return P.is_ip ? launch_kernel(ivfpq_scan_m64_kernel<true, 384, MINB, PER>, P, 2, 384, &conf[0], st)
                 : launch_kernel(ivfpq_scan_m64_kernel<false, 384, MINB, PER>, P, 2, 384, &conf[1], st);
}

cudaError_t launch_ivfpq_scan(const ScanParams &P, int mode, cudaStream_t st) {
  static size_t conf[4] = {0, 0, 0, 0};
  if (mode == 2) return P.cap <= 4 * 384 ? launch_m64<4>(P, st) : launch_m64<16>(P, st);
  if (mode == 1) return cudaErrorInvalidValue;  // M = 32 runs the persistent kernel (launch_ivfpq_scan_v3)
  if (P.cap > 4 * SCAN_THREADS)
    return P.is_ip ? launch_kernel(ivfpq_scan_generic_kernel<true, 16>, P, 0, SCAN_THREADS, &conf[0], st)
                   : launch_kernel(ivfpq_scan_generic_kernel<false, 16>, P, 0, SCAN_THREADS, &conf[1], st);
  return P.is_ip ? launch_kernel(ivfpq_scan_generic_kernel<true, 4>, P, 0, SCAN_THREADS, &conf[2], st)
                 : launch_kernel(ivfpq_scan_generic_kernel<false, 4>, P, 0, SCAN_THREADS, &conf[3], st);
}

}  // namespace gb
