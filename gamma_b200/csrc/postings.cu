// K0 — device-side maintenance of the posting mirror: append / overwrite a posting in the
// blocked layout the scan reads, compute its query-independent L2 term t(p), read a list
// back in the reference's AoS form, and build the per-search validity bitmap.
// Reference being mirrored (file:line):
//   RealTimeMemData::AddKeys / Update      realtime/realtime_mem_data.cc:264-327
//   RTInvertBucketData layout (AoS)        realtime/realtime_mem_data.h:29-68
//   IndexIVFPQ::precompute_table (term 2)  faiss IndexIVFPQ.cpp:411-453
//   GammaSearchCondition::IsValid          common/gamma_common_data.h:99-108
//   RangeQueryResult::Has / bitmap::test   table/range_query_result.h:53-67, util/bitmap.cc:25-27
#include "common.cuh"
#include "kernels.h"

namespace gb {

__device__ __forceinline__ long long code_byte_addr(long long blk_base, int i, int b, int chunk) {
  return blk_base + ((long long)(b / chunk) * 32 + i) * chunk + (b % chunk);
}

__global__ void append_kernel(AppendParams P) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.n) return;
  const int M = P.M;
  const int list = P.list_no[t];
  const int pos = P.pos[t];
  const uint8_t *code = P.codes_aos + t * M;
  const long long off = P.list_off[list];
  const long long pidx = off + pos;
  const int i = pos & 31;
  const long long blk_base = (off + (pos & ~31)) * (long long)M;
  // codes: compose 4-byte words of the stored order
  for (int w = 0; w < M; w += 4) {
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int s = w + j;
      int src = (P.layout == LAYOUT_PLAIN) ? s : ((i + s) % M);  // rotated layouts: M = 32 or 64
      word |= (uint32_t)code[src] << (8 * j);
    }
    *reinterpret_cast<uint32_t *>(P.codes + code_byte_addr(blk_base, i, w, P.chunk)) = word;
  }
  // t(p) = SUM_m |cb[m][c_m]|^2 + 2 * centroid[list]_m . cb[m][c_m]
  if (P.norms) {
    const float *cen = P.centroids + (size_t)list * P.d;
    double acc = 0.0;
    for (int m = 0; m < M; m++) {
      const float *cb = P.pq + ((size_t)m * 256 + code[m]) * P.dsub;
      const float *cm = cen + m * P.dsub;
      float s = 0.f;
      for (int j = 0; j < P.dsub; j++) {
        float v = __ldg(cb + j);
        s = fmaf(v, v, s);
        s = fmaf(2.f * __ldg(cm + j), v, s);
      }
      acc += (double)s;
    }
    P.norms[pidx] = (float)acc;
  }
  P.ids[pidx] = P.vid[t];
}

cudaError_t launch_append(const AppendParams &P, cudaStream_t st) {
  if (P.n <= 0) return cudaSuccess;
  int threads = 128;
  long long blocks = (P.n + threads - 1) / threads;
  append_kernel<<<(unsigned)blocks, threads, 0, st>>>(P);
  return cudaGetLastError();
}

__global__ void fill_i32_kernel(int *p, long long n, int v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
cudaError_t launch_fill_i32(int *p, long long n, int v, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fill_i32_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, n, v);
  return cudaGetLastError();
}

__global__ void gather_list_kernel(const uint8_t *codes, const int *ids, long long off, int len, int M,
                                   int chunk, int layout, uint8_t *out_codes, int *out_ids) {
  int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= len) return;
  int i = pos & 31;
  long long blk_base = (off + (pos & ~31)) * (long long)M;
  for (int s = 0; s < M; s++) {
    uint8_t v = codes[code_byte_addr(blk_base, i, s, chunk)];
    int dst = (layout == LAYOUT_PLAIN) ? s : ((i + s) % M);
    out_codes[(size_t)pos * M + dst] = v;
  }
  out_ids[pos] = ids[off + pos];
}
cudaError_t launch_gather_list(const uint8_t *codes, const int *ids, long long off, int len, int M,
                               int chunk, int layout, uint8_t *out_codes, int *out_ids, cudaStream_t st) {
  if (len <= 0) return cudaSuccess;
  gather_list_kernel<<<(len + 127) / 128, 128, 0, st>>>(codes, ids, off, len, M, chunk, layout, out_codes, out_ids);
  return cudaGetLastError();
}

// valid[w] = live[w] & AND_f has_f(doc)   for the 32 docs of word w  (live: bit = 1 <=> the doc is NOT deleted; words the
// live bitmap does not cover are all ones)
__global__ void build_valid_kernel(const uint32_t *live, long long live_words, const DevRangeFilter *filters,
                                   int n_filters, uint32_t *valid, long long nwords) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  uint32_t v = (live && w < live_words) ? live[w] : 0xffffffffu;
  for (int f = 0; f < n_filters && v; f++) {
    DevRangeFilter rf = filters[f];
    uint32_t m = 0;
    for (int b = 0; b < 32; b++) {
      long long doc = w * 32 + b;
      bool in;
      if (doc < rf.min_doc || doc > rf.max_doc) {
        in = rf.not_in != 0;
      } else {
        long long r = doc - rf.min_aligned;
        bool bit = (rf.bitmap[r >> 3] >> (r & 7)) & 1;
        in = rf.not_in ? !bit : bit;
      }
      m |= (in ? 1u : 0u) << b;
    }
    v &= m;
  }
  valid[w] = v;
}
cudaError_t launch_build_valid(const uint32_t *live, long long live_bits, const DevRangeFilter *filters,
                               int n_filters, uint32_t *valid, long long nbits, cudaStream_t st) {
  long long nwords = (nbits + 31) / 32;
  if (nwords <= 0) return cudaSuccess;
  build_valid_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(live, (live_bits + 31) / 32, filters, n_filters,
                                                                        valid, nwords);
  return cudaGetLastError();
}

// ---- publication of list extents (RealTimeMemData's retrieve_idx_pos_ bump, realtime_mem_data.cc:299-301, and its
// copy-swap of a grown bucket, :426-474).  Readers (probe setup kernels) load len with acquire semantics and THEN off;
// the writer stores off and THEN len with release semantics: a reader can pair a new off with an old len (the new
// region holds everything the old one did) but never an old off with a new len.
__global__ void publish_lists_kernel(const int *lists, const long long *offs, const int *lens, int n, long long *d_off,
                                     int *d_len) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = lists[i];
  d_off[l] = offs[i];
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(d_len + l), "r"(lens[i]) : "memory");
}
cudaError_t launch_publish_lists(const int *lists, const long long *offs, const int *lens, int n, long long *d_off,
                                 int *d_len, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  publish_lists_kernel<<<(n + 255) / 256, 256, 0, st>>>(lists, offs, lens, n, d_off, d_len);
  return cudaGetLastError();
}

// words[idx[i]] = val[i]  (touched words of the live-docs bitmap after BitmapManager::Set / Unset)
__global__ void scatter_words_kernel(const long long *idx, const uint32_t *val, int n, uint32_t *words) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) words[idx[i]] = val[i];
}
cudaError_t launch_scatter_words(const long long *idx, const uint32_t *val, int n, uint32_t *words, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  scatter_words_kernel<<<(n + 255) / 256, 256, 0, st>>>(idx, val, n, words);
  return cudaGetLastError();
}

// ---- compaction of posting lists on the device (RealTimeMemData::CompactBucket / CompactOne,
// realtime_mem_data.cc:98-112, 354-424): drop postings that were moved away (kDelIdxMask -> negative id here) or whose
// doc is deleted in the bitmap, keep the order of the survivors.  One CTA per list.
//   pass 0 (dst_codes == nullptr): new_len[l] = survivors of list l;
//   pass 1: survivors are written, in order, to the region starting at new_off[l]; ids behind them are set to -1 up to
//           new_cap[l].  In the rotated layouts a posting's stored bytes depend on its lane (position mod 32), so the
//           code bytes are re-rotated for the new position.
__device__ __forceinline__ bool posting_survives(int id, const uint32_t *live, long long live_bits) {
  if (id < 0) return false;
  if (live && (long long)id < live_bits) return (live[id >> 5] >> (id & 31)) & 1u;
  return true;
}
__global__ void __launch_bounds__(256) compact_lists_kernel(CompactParams P) {
  const int l = P.lists ? P.lists[blockIdx.x] : blockIdx.x;
  const long long off = P.list_off[l];
  const int len = P.list_len[l];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int warp_cnt[8];
  __shared__ int run_base;
  if (tid == 0) run_base = 0;
  __syncthreads();
  const int M = P.M;
  for (int p0 = 0; p0 < len; p0 += 256) {
    const int pos = p0 + tid;
    int id = -1;
    if (pos < len) id = P.ids[off + pos];
    const bool keep = pos < len && posting_survives(id, P.live, P.live_bits);
    const unsigned m = __ballot_sync(GB_FULL, keep);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int before = run_base;
    for (int w = 0; w < warp; w++) before += warp_cnt[w];
    const int npos = before + __popc(m & ((1u << lane) - 1u));
    if (keep && P.dst_codes) {
      const long long noff = P.new_off[blockIdx.x];
      P.dst_ids[noff + npos] = id;
      if (P.dst_norms) P.dst_norms[noff + npos] = P.norms[off + pos];
      const int i = pos & 31, ni = npos & 31;
      const long long sb = (off + (pos & ~31)) * (long long)M, db = (noff + (npos & ~31)) * (long long)M;
      for (int s = 0; s < M; s++) {  // destination stored byte s = logical byte (ni + s) % M (rotated) or s (plain)
        const int logical = P.layout == LAYOUT_PLAIN ? s : (ni + s) % M;
        const int src_s = P.layout == LAYOUT_PLAIN ? logical : (logical - i + M) % M;
        P.dst_codes[code_byte_addr(db, ni, s, P.chunk)] = P.codes[code_byte_addr(sb, i, src_s, P.chunk)];
      }
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 8; w++) t += warp_cnt[w];
      run_base += t;
    }
    __syncthreads();
  }
  const int nl = run_base;
  if (!P.dst_codes) {
    if (tid == 0) P.new_len[blockIdx.x] = nl;
    return;
  }
  const long long noff = P.new_off[blockIdx.x];
  for (int p = nl + tid; p < P.new_cap[blockIdx.x]; p += 256) P.dst_ids[noff + p] = -1;
}
cudaError_t launch_compact_lists(const CompactParams &P, int n_lists, cudaStream_t st) {
  if (n_lists <= 0) return cudaSuccess;
  compact_lists_kernel<<<n_lists, 256, 0, st>>>(P);
  return cudaGetLastError();
}

__global__ void row_norms_kernel(const float *x, int rows, int d, float *out) {
  int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  int lane = threadIdx.x & 31;
  const float *p = x + (size_t)r * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s = fmaf(p[i], p[i], s);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(GB_FULL, s, o);
  if (lane == 0) out[r] = s;
}
// |x|^2 (same summation order as row_norms_kernel) and x - tf32(x) of every row in ONE pass: the two operands the
// tensor-core distance producer needs next to the rows themselves (one launch instead of two on the search path)
__global__ void rows_prep_kernel(const float *x, int rows, int d, float *norms, float *small, int *zero_words, int n_zero,
                                 unsigned long long *zero_u64) {
  // optional: control words of a later kernel of the same search, zeroed here instead of by separate memset nodes
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += gridDim.x * blockDim.x) zero_words[i] = 0;
  if (zero_u64 && blockIdx.x == 0 && threadIdx.x == 0) *zero_u64 = 0ull;
  int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  int lane = threadIdx.x & 31;
  const float *p = x + (size_t)r * d;
  float *q = small + (size_t)r * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float v = p[i];
    s = fmaf(v, v, s);
    q[i] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(GB_FULL, s, o);
  if (lane == 0) norms[r] = s;
}
cudaError_t launch_rows_prep(const float *x, int rows, int d, float *norms, float *small, cudaStream_t st, int *zero_words,
                             int n_zero, unsigned long long *zero_u64) {
  if (rows <= 0) return cudaSuccess;
  rows_prep_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, rows, d, norms, small, zero_words, zero_words ? n_zero : 0, zero_u64);
  return cudaGetLastError();
}

// y[r][o] = b[o] + sum_k x[r][k] * At[k][o]   (At = A transposed: [d_in][d_out]; OPQMatrix::apply, fp32 FMA chain over k)
__global__ void __launch_bounds__(256) linear_apply_kernel(const float *__restrict__ x, int x_stride, int d_in,
                                                           const float *__restrict__ At, const float *__restrict__ b,
                                                           int d_out, float *__restrict__ y) {
  extern __shared__ float xs[];
  const int r = blockIdx.x;
  for (int k = threadIdx.x; k < d_in; k += blockDim.x) xs[k] = k < x_stride ? x[(size_t)r * x_stride + k] : 0.f;
  __syncthreads();
  for (int o = threadIdx.x; o < d_out; o += blockDim.x) {
    float s = b ? b[o] : 0.f;
    for (int k = 0; k < d_in; k++) s = fmaf(xs[k], __ldg(At + (size_t)k * d_out + o), s);
    y[(size_t)r * d_out + o] = s;
  }
}
cudaError_t launch_linear_apply(const float *x, int x_stride, int rows, int d_in, const float *At, const float *b,
                                int d_out, float *y, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  const int threads = d_out >= 256 ? 256 : ((d_out + 31) & ~31);
  linear_apply_kernel<<<rows, threads, (size_t)d_in * sizeof(float), st>>>(x, x_stride, d_in, At, b, d_out, y);
  return cudaGetLastError();
}

cudaError_t launch_row_norms(const float *x, int rows, int d, float *out, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  row_norms_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, rows, d, out);
  return cudaGetLastError();
}

}  // namespace gb
