// Host-visible launchers of the gamma_b200 kernels (internal; the public surface is
// include/gamma_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gb {

typedef unsigned long long u64;

// scan-order field of a candidate key: (probe_rank << 21) | position_in_list
#define GB_SEQ_POS_BITS 21  // bucket_max_size default 1,280,000 < 2^21 (gamma_index_ivfpq.h:706)
#define GB_SEQ_POS_MASK ((1u << GB_SEQ_POS_BITS) - 1u)

// posting-mirror layouts (how a 32-posting block of codes is laid out in HBM)
//   0: chunk-major, chunk = largest of {16,8,4} dividing M: byte b of posting i at
//      ((b/chunk)*32 + i)*chunk + b%chunk
//   1: M == 32, chunk 16, bytes pre-rotated per lane: stored byte s of posting i = code[(i + s) % 32]
//   2: M == 64, chunk 16, bytes pre-rotated per lane: stored byte s of posting i = code[(i + s) % 64]  (opt-in)
enum { LAYOUT_PLAIN = 0, LAYOUT_M32_ROT = 1, LAYOUT_M64_ROT = 2 };

struct ScanParams {
  const float *xq;          // [n][d]
  const int *keys;          // [n][nprobe] probed lists, ascending coarse distance, -1 = none
  const float *coarse_dis;  // [n][nprobe]
  const float *centroids;   // [nlist][d]
  const float *pq_t;        // [256][M][dsub] code-major PQ codebook
  const float *lut_g;       // M = 32 kernel: [n][256][64] per-query tables from launch_lut_build_m32
  const uint8_t *codes;     // posting pool (layout above)
  const int *ids;           // [pool] vid, -1 = padding / moved (kDelIdxMask)
  const float *norms;       // [pool] t(p) (L2 only)
  const long long *list_off;// [nlist] first posting of the list (multiple of 32)
  const int *list_len;      // [nlist]
  const uint32_t *valid;    // validity bitmap or nullptr (everything valid)
  long long valid_bits;     // docs the bitmap covers; ids beyond it (appended after the search began) are skipped
  u64 *cand;                // [n][S][R] surviving keys (unsorted), GB_KEY_MAX padded
  unsigned long long *scanned;  // += postings walked (may be nullptr)
  int n, d, M, dsub, nlist, nprobe, S, R, cap, chunk, max_np_s, is_ip;
  int m32_threads;          // tuning: CTA size of the M = 32 kernel (256 / 320 / 384)
  int variant;              // 3 = persistent M = 32 kernel (ivfpq_scan_v3.cu), 2 = M = 64 kernel, 0 = generic
  int pf_blocks;            // M = 64 loop: L2 prefetch distance in 32-posting blocks (0 = off)
  unsigned char *probe_g;   // v2: [items][scan_probe_bytes(max_np_s)] from launch_probe_setup
  int n_items;              // v2: > 0: grid = n_items work items; S is then the row count of cand per query.  The plan is
  int n_full, s_tail;       //     positional: query q < n_full is one item, the others s_tail items each
  // v3 (persistent kernel, ivfpq_scan_v3.cu): control words, zeroed before every launch.  S = candidate rows per query.
  int *v3_next_q;           // [1] next query without an owner CTA
  int *v3_claim;            // [n] next unclaimed item of the query (warps of every CTA working on it claim here)
  int *v3_rows;             // [n] candidate rows handed out for the query (>= S: no row left)
  int ch_blocks;            // 32-posting blocks per item
  int v3_zero;              // always 0 (keeps the claim address opaque to the compiler, see scan_loop_m32_v3)
  int help_min;             // an idle CTA joins a running query that still has >= help_min unclaimed items ...
  int help_window;          // ... looking at the last help_window queries
  int v3_flags;             // bit 0: next query fetched inside the scan, bit 1: its tables requested before the final select,
                            // bit 2: approximate in-loop prunes
  int v3_max_items;         // capacity of the per-query item table (host bound: nprobe x items of the longest list)
  int v3_tma;               // posting ring fed by bulk copies (cp.async.bulk, one elected lane) instead of per-lane cp.async
};
// v3 launchers (ivfpq_scan_v3.cu)
size_t scan_v3_probe_bytes(int nprobe, int max_items);
size_t scan_v3_smem_bytes_for(int nprobe, int max_items, int cap, int threads);
int scan_v3_ctas_per_sm(int threads, int cap);
cudaError_t launch_probe_setup_v3(const ScanParams &P, cudaStream_t st);
cudaError_t launch_ivfpq_scan_v3(const ScanParams &P, int grid, cudaStream_t st);
size_t scan_probe_bytes_host(int max_np_s);
cudaError_t launch_probe_setup(const ScanParams &P, cudaStream_t st);
size_t scan_smem_bytes(const ScanParams &P, int mode);
int scan_buffer_cap(int R);
cudaError_t launch_ivfpq_scan(const ScanParams &P, int mode, cudaStream_t st);
cudaError_t launch_lut_build_m32(const float *xq, const float *pq_t, float *lut_g, int n, int d, int dsub, int is_ip,
                                 cudaStream_t st);
cudaError_t launch_lut_build_m64(const float *xq, const float *pq_t, float *lut_g, int n, int d, int dsub, int is_ip,
                                 cudaStream_t st);  // [n][256][96]

// K0 — device-side append into the posting mirror (+ t(p) for L2)
struct AppendParams {
  const int *list_no;       // [n]
  const int *pos;           // [n] position inside the list
  const int *vid;           // [n]
  const uint8_t *codes_aos; // [n][M]
  const float *centroids;   // [nlist][d]
  const float *pq;          // [M][256][dsub]
  const long long *list_off;
  uint8_t *codes;
  int *ids;
  float *norms;
  int n, d, M, dsub, chunk, layout;
};
cudaError_t launch_append(const AppendParams &P, cudaStream_t st);
cudaError_t launch_fill_i32(int *p, long long n, int v, cudaStream_t st);
// read a list back in reference AoS form (test hook)
cudaError_t launch_gather_list(const uint8_t *codes, const int *ids, long long off, int len, int M,
                               int chunk, int layout, uint8_t *out_codes, int *out_ids, cudaStream_t st);

// publish list extents after an append / relocation / compaction (off, then len with release semantics)
cudaError_t launch_publish_lists(const int *lists, const long long *offs, const int *lens, int n, long long *d_off,
                                 int *d_len, cudaStream_t st);
cudaError_t launch_scatter_words(const long long *idx, const uint32_t *val, int n, uint32_t *words, cudaStream_t st);
// device-side list compaction (one CTA per entry of `lists`, or per list when lists == nullptr)
struct CompactParams {
  const int *lists;          // [n_lists] or nullptr = all lists 0..n_lists-1
  const long long *list_off; // current extents
  const int *list_len;
  const uint8_t *codes;      // source pools
  const int *ids;
  const float *norms;        // may be nullptr (InnerProduct)
  const uint32_t *live;      // live-docs bitmap (bit = 1: not deleted) or nullptr
  long long live_bits;
  int *new_len;              // pass 0 out: [n_lists] survivors per list
  const long long *new_off;  // pass 1 in: [n_lists] destination region of each list
  const int *new_cap;        //            [n_lists] its capacity (ids behind the survivors are set to -1)
  uint8_t *dst_codes;        // pass 1: destination pools (nullptr = pass 0, count only)
  int *dst_ids;
  float *dst_norms;
  int M, chunk, layout;
};
cudaError_t launch_compact_lists(const CompactParams &P, int n_lists, cudaStream_t st);

// K1 — coarse quantiser: dist[n][nlist] = |q|^2 + |c|^2 - 2 q.c (clamped at 0), then top-nprobe
cudaError_t launch_row_norms(const float *x, int rows, int d, float *out, cudaStream_t st);
// OPQ pre-transform: y = A x + b for every row (At = A transposed, [d_in][d_out]; x rows x_stride wide, zero beyond)
cudaError_t launch_linear_apply(const float *x, int x_stride, int rows, int d_in, const float *At, const float *b,
                                int d_out, float *y, cudaStream_t st);
// |x|^2 and x - tf32(x) of every row in one pass (the query side of the tensor-core distance producer)
// permission for a launch with `smem` bytes of dynamic shared memory (> 48 KB needs the opt-in attribute): grows the
// attribute of (func, current device) monotonically, safe against concurrent launches with other sizes (launch.cu)
cudaError_t ensure_dynamic_smem(const void *func, size_t smem);
template <typename K>
inline cudaError_t ensure_dynamic_smem(K *kernel, size_t smem) {
  return ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), smem);
}

// zero_words / zero_u64 (optional): control words of a later kernel of the same search, zeroed by this launch
cudaError_t launch_rows_prep(const float *x, int rows, int d, float *norms, float *small, cudaStream_t st,
                             int *zero_words = nullptr, int n_zero = 0, unsigned long long *zero_u64 = nullptr);
cudaError_t launch_coarse_dist(const float *xq, const float *xq_norm, const float *cent,
                               const float *cent_norm, int n, int nlist, int d, float *dist,
                               cudaStream_t st);
// tensor-core distance producer (tc_gemm.cu): out[M][ldo] = L2^2 (l2=1) or inner product of a (M x K) vs b (N x K)
cudaError_t launch_tf32_residual(const float *x, float *small, size_t n, cudaStream_t st);
// cmin (optional, l2 only): [M][cmin_pitch] minimum of every 32-column chunk of out, cmin_pitch >= tc_gemm_cmin_pitch(N)
cudaError_t launch_tc_gemm(const float *a, const float *a_small, const float *a_norm, const float *b,
                           const float *b_small, const float *b_norm, int M, int N, int K, float *out, int ldo, int l2,
                           float *cmin, int cmin_pitch, cudaStream_t st);
inline int tc_gemm_cmin_pitch(int N) { return ((N + 127) / 128) * 4; }
cudaError_t launch_coarse_select(const float *dist, int n, int nlist, int nprobe, int *keys,
                                 float *coarse_dis, cudaStream_t st);
// select that starts from the GEMM's chunk minima: reads ~nprobe chunks of the row instead of all of it
bool coarse_select_cmin_usable(int nlist, int nprobe);
cudaError_t launch_coarse_select_cmin(const float *dist, const float *cmin, int cmin_pitch, int n, int nlist, int nprobe,
                                      int *keys, float *coarse_dis, cudaStream_t st);

// K5 — encode half of Add (encode.cu): residual against the assigned centroid + PQ argmin per sub-quantiser, in faiss'
// own arithmetic.  x rows are x_stride floats wide (columns beyond it read as zero), keys from the coarse stage.
cudaError_t launch_pq_encode(const float *x, int x_stride, const int *keys, const float *centroids, const float *pq,
                             long long n, int d, int M, int dsub, int by_residual, uint8_t *codes, cudaStream_t st);

// K7 — IVFFLAT scan (ivfflat.cu): exact distances over the probed lists (vids only; vectors come from the raw store)
struct IvfFlatParams {
  const float *xq;          // [n][d]
  const int *keys;          // [n][nprobe]
  const long long *list_off;
  const int *list_len;
  const int *ids;
  const float *raw;         // [nraw][d]
  long long nraw;
  const uint32_t *valid;    // or nullptr
  long long valid_bits;
  u64 *cand;                // [n][S][R]
  unsigned long long *scanned;
  int n, d, nlist, nprobe, S, R, cap, is_ip;
  float min_score, max_score;
};
int ivfflat_buffer_cap(int R);
cudaError_t launch_ivfflat_scan(const IvfFlatParams &P, cudaStream_t st);

// K3 — merge the per-split survivors, optional exact re-rank, score window, top-k
// Multi-GPU result sink of the final kernel of a search (comm.cu): the kernel stores every query's k results into each
// peer's result window as well (peer memory over NVLink), and its last CTA raises this rank's flag there — the exchange
// rides on the kernel that produces the results.
constexpr int GB_MAX_PEERS = 15;
struct PeerSink {
  int n_peers;                        // 0 = none
  float *dist[GB_MAX_PEERS];          // [n][k] in peer p's window
  long long *ids[GB_MAX_PEERS];       // [n][k]
  uint32_t *flag[GB_MAX_PEERS];       // this rank's flag in peer p's window (release store of `epoch`)
  const uint32_t *wait[GB_MAX_PEERS]; // peer p's flag in MY window to wait for (>= wait_epoch), or nullptr
  uint32_t epoch, wait_epoch;
  unsigned int *done;                 // CTA counter (zero between launches)
  unsigned int *err;                  // set to 1 + p when peer p did not arrive in time
  int peer_rank[GB_MAX_PEERS];
};

struct RerankParams {
  const u64 *cand;          // [n][S][R]
  const int *keys;          // [n][nprobe]
  const long long *list_off;
  const int *ids;
  const float *xq;          // [n][raw_d] (raw query, rerank uses raw_d)
  const float *raw;         // [nraw][raw_d]
  long long nraw;
  float *out_dist;          // [n][k]
  long long *out_ids;       // [n][k]
  int n, S, R, k, nprobe, raw_d, xq_stride, has_rank, is_ip;
  float min_score, max_score;
  int stage_rows, stage_off;  // set by launch_rerank: raw rows staged in shared memory per chunk / their offset (0 = off)
  int no_stage;             // tuning: 1 = the L2-prefetch + load-on-use path
  const int *nsplit;        // optional [n]: candidate rows the query really has (<= S); nullptr = S ...
  int n_full;               // ... or, when > 0 (positional plan), 1 row for q < n_full and S rows otherwise
  PeerSink sink;            // multi-GPU: also store the results into the peers' windows (n_peers = 0: off)
};
cudaError_t launch_rerank(const RerankParams &P, cudaStream_t st);

// K4 — flat scan over the raw vectors
struct FlatParams {
  const float *xq;          // [n][d]
  const float *raw;         // [N][d]
  const uint32_t *valid;    // or nullptr
  long long N;
  float *out_dist;
  long long *out_ids;
  int n, d, k, is_ip;
  float min_score, max_score;
  u64 *scratch;             // [n][nsplit][kpad]
  int nsplit;
};
cudaError_t launch_flat_exact(const FlatParams &P, cudaStream_t st);
int flat_exact_splits(long long N, int n);

// K4t — batched flat through the tensor cores (flat_tc.cu)
int flat_tc_candidates(int k);
cudaError_t launch_flat_chunk_select(const float *dist, int ldo, int nc, long long chunk_base, const uint32_t *valid,
                                     float lo, float hi, int Kp, int first, u64 *state, int n, int is_ip,
                                     cudaStream_t st);
cudaError_t launch_flat_rescore(const u64 *state, int Kp, const float *xq, const float *raw, int n, int d,
                                float min_score, float max_score, int k, int is_ip, float *out_d, long long *out_i,
                                cudaStream_t st);

// test hook: BlockTopR selection on caller-provided keys (single CTA)
cudaError_t launch_select_selftest(const u64 *keys, int n, int R, int cap, int batch, int threads, u64 *out, int *out_n,
                                   cudaStream_t st);

// validity bitmap = live (NOT deleted) AND all range filters
struct DevRangeFilter {
  const uint8_t *bitmap;  // device bytes
  int min_doc, max_doc, min_aligned, not_in;
};
cudaError_t launch_build_valid(const uint32_t *live, long long live_bits, const DevRangeFilter *filters,
                               int n_filters, uint32_t *valid, long long nbits, cudaStream_t st);

}  // namespace gb

// internal (capi.cu, used by comm.cu): gb200_ivfpq_search_dev whose final kernel also stores the results into the peers'
// windows and raises / awaits the exchange flags (sink); *sink_used = 0 when the search took a path without that kernel
struct gb200_index;
struct gb200_search_params;
int gb_ivfpq_search_dev_sink(gb200_index *ix, int n, const float *xq_dev, int k, const gb200_search_params *sp, float *D_dev,
                             int64_t *I_dev, void *stream, const gb::PeerSink *sink, int *sink_used);
