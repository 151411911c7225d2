// K7 — IVFFLAT scan: exact L2 / inner-product distances over the probed inverted lists, streaming top-k.
// Replaces GammaIVFFlatScanner1::scan_codes (index/impl/gamma_index_ivfflat.h:59-84) inside
// GammaIndexIVFFlat::search_preassigned (index/impl/gamma_index_ivfflat.cc:423-560): kDelIdxMask test, IsValid
// (deleted bitmap AND range filters), fvec_L2sqr / fvec_inner_product, IsSimilarScoreValid window, heap with the strict
// compare (first-scanned posting wins among equal distances).
//
// The device keeps ONE copy of every vector (the raw store, indexed by vid) and the lists hold vids only, so a posting
// costs a gathered raw_d * 4-byte row read instead of a contiguous one: HBM-bound like the reference's list walk, at
// the DRAM efficiency of row-sized random reads.  Distances use the AVX summation order of the CPU kernels
// (exact_distance_octet), so they are bit-identical to the reference's.
//
// One CTA per (query, split), probes dealt round-robin over the splits; 8 lanes per posting, 32 postings per pass.
// Survivors go to cand[q][split][0..R) in the scan kernels' key format and are merged by rerank_kernel's no-rank branch.
#include "scan_common.cuh"

namespace gb {

constexpr int IF_THREADS = 256;
constexpr int IF_OCT = IF_THREADS / 8;  // postings per pass
constexpr int IF_CHECK = 4;             // passes between overflow checks

template <bool IP, int PER>
__global__ void __launch_bounds__(IF_THREADS) ivfflat_scan_kernel(IvfFlatParams P) {
  extern __shared__ __align__(16) unsigned char if_smem[];
  u64 *buf = reinterpret_cast<u64 *>(if_smem);
  int *misc = reinterpret_cast<int *>(if_smem + (size_t)P.cap * sizeof(u64));
  float *qs = reinterpret_cast<float *>(misc + 4 + 64);
  BlockTopR topr;
  topr.buf = buf;
  topr.tau = reinterpret_cast<u64 *>(misc);
  topr.cnt = misc + 2;
  topr.warp_part = misc + 4;
  topr.cap = P.cap;
  topr.R = P.R;
  topr.init_collective();
  const int q = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  const int sub = tid & 7, oct = tid >> 3;
  for (int i = tid; i < P.d; i += IF_THREADS) qs[i] = P.xq[(size_t)q * P.d + i];
  __syncthreads();
  const uint32_t valid_lim = (uint32_t)(P.valid_bits < 0x7fffffffLL ? P.valid_bits : 0x7fffffffLL);
  const int prune_limit = P.cap - IF_CHECK * IF_OCT;
  unsigned long long walked = 0;
  int pass_no = 0;
  for (int j = split; j < P.nprobe; j += P.S) {
    const int key = P.keys[(size_t)q * P.nprobe + j];
    if (key < 0 || key >= P.nlist) continue;  // uniform over the CTA
    long long off;
    int len;
    load_list_extent(P.list_off, P.list_len, key, off, len);
    walked += (unsigned long long)len;
    for (int base = 0; base < len; base += IF_OCT) {
      const int pos = base + oct;
      int id = pos < len ? ldg_nc_s32(P.ids + off + pos) : -1;
      bool alive = id >= 0 && (long long)id < P.nraw;  // negative: moved away (kDelIdxMask) or padding
      if (alive && P.valid) alive = (uint32_t)id < valid_lim && bitmap_test(P.valid, id);
      const float *y = P.raw + (size_t)(alive ? id : 0) * P.d;
      const float dis = exact_distance_octet<IP>(qs, y, alive ? P.d : 0, sub);
      const u64 k64 = ((u64)dist_to_key32<IP>(dis) << 32) | (((uint32_t)j << GB_SEQ_POS_BITS) | (uint32_t)pos);
      const bool pass = alive && sub == 0 && dis >= P.min_score && dis <= P.max_score && k64 < topr.threshold();
      topr.append_warp(pass, k64);
      if (++pass_no == IF_CHECK) {
        pass_no = 0;
        const int over = *((volatile int *)topr.cnt) > prune_limit;
        if (__syncthreads_or(over)) topr.prune_collective<PER>();
      }
    }
  }
  topr.prune_collective<PER>();
  const int n_out = min(*((volatile int *)topr.cnt), P.R);
  u64 *out = P.cand + ((size_t)q * P.S + split) * P.R;
  for (int i = tid; i < P.R; i += IF_THREADS) out[i] = i < n_out ? buf[i] : GB_KEY_MAX;
  if (P.scanned && tid == 0 && walked) atomicAdd(P.scanned, walked);
}

int ivfflat_buffer_cap(int R) {
  int cap = 1024;
  while (cap < R + 2 * IF_CHECK * IF_OCT) cap <<= 1;
  return cap;
}

cudaError_t launch_ivfflat_scan(const IvfFlatParams &P, cudaStream_t st) {
  if (P.n <= 0) return cudaSuccess;
  if (P.cap > 16 * IF_THREADS) return cudaErrorInvalidValue;
  const size_t smem = (size_t)P.cap * sizeof(u64) + (4 + 64) * sizeof(int) + (size_t)P.d * sizeof(float);
  dim3 grid(P.S, P.n);
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = ensure_dynamic_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, IF_THREADS, smem, st>>>(P);
    return cudaGetLastError();
  };
  const bool big = P.cap > 4 * IF_THREADS;
  if (P.is_ip) return big ? go(ivfflat_scan_kernel<true, 16>) : go(ivfflat_scan_kernel<true, 4>);
  return big ? go(ivfflat_scan_kernel<false, 16>) : go(ivfflat_scan_kernel<false, 4>);
}

}  // namespace gb
