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

// valid[w] = ~deleted[w] & AND_f has_f(doc)   for the 32 docs of word w
__global__ void build_valid_kernel(const uint32_t *deleted, long long deleted_words, const DevRangeFilter *filters,
                                   int n_filters, uint32_t *valid, long long nwords) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  uint32_t v = (deleted && w < deleted_words) ? ~deleted[w] : 0xffffffffu;
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
cudaError_t launch_build_valid(const uint32_t *deleted, long long deleted_bits, const DevRangeFilter *filters,
                               int n_filters, uint32_t *valid, long long nbits, cudaStream_t st) {
  long long nwords = (nbits + 31) / 32;
  if (nwords <= 0) return cudaSuccess;
  build_valid_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(deleted, (deleted_bits + 31) / 32, filters,
                                                                        n_filters, valid, nwords);
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
cudaError_t launch_row_norms(const float *x, int rows, int d, float *out, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  row_norms_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, rows, d, out);
  return cudaGetLastError();
}

}  // namespace gb
