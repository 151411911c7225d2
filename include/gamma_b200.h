/*
 * gamma_b200.h — C ABI of the B200-native Gamma search hot path.
 *
 * This is the drop-in boundary BELOW the reference's RetrievalModel plugin
 * interface (/root/reference/index/retrieval_model.h:218-310): the C++ plugin
 * classes in gamma_b200/plugin/ ("B200IVFPQ", "B200FLAT", registered with
 * REGISTER_MODEL) are thin hosts that forward to these entry points, exactly
 * where the reference's own models call into faiss.  Plain pointers and sizes
 * only — no C++/torch types — so Go (cgo), Python (ctypes) or C++ can bind it.
 *
 * Conventions (reference: search/error_code.h:17-25, retrieval_model.h:228-301):
 *   - every function returns 0 on success, a negative GB200_E* code on failure;
 *     no exception ever crosses this boundary;
 *   - pointers are HOST pointers unless the parameter name ends in _dev;
 *   - output buffers are caller-allocated; unfilled result slots are id = -1 and
 *     distance = FLT_MAX (L2) / -FLT_MAX (InnerProduct), the neutral heap value
 *     the reference leaves there (faiss utils/ordered_key_value.h:49-69);
 *   - there is NO CPU fallback: if no sm_100 device is usable, create fails.
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef GAMMA_B200_H_
#define GAMMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB200_OK 0
#define GB200_EINVAL (-1)      /* bad argument (reference: -1 "bad input")            */
#define GB200_ENOTTRAINED (-2) /* quantizers not set (reference Search returns -2)     */
#define GB200_ECUDA (-3)       /* CUDA runtime error; see gb200_last_error()           */
#define GB200_ENOMEM (-4)      /* device/host allocation failed                        */
#define GB200_EUNSUPPORTED (-5)/* parameter outside what the kernels implement         */

/* DistanceComputeType values of index/retrieval_model.h:20 */
#define GB200_METRIC_INNER_PRODUCT 0
#define GB200_METRIC_L2 1

typedef struct gb200_index gb200_index; /* opaque: one RetrievalModel's device state */

/* Model parameters: IVFPQModelParams (index/impl/gamma_index_ivfpq.h:675-887) reduced
 * to what the search path needs.  nbits_per_idx must be 8, by_residual is always true,
 * the coarse quantizer is always IndexFlatL2 (gamma_index_ivfpq.cc:147,179).          */
typedef struct gb200_ivfpq_params {
  int device;      /* CUDA device ordinal                                              */
  int d;           /* index dimension (d_ after support_indivisible_nsubvector pad)    */
  int raw_d;       /* raw vector dimension (re-rank / flat use this)                   */
  int nlist;       /* ncentroids                                                       */
  int nsubvector;  /* M                                                                */
  int nbits;       /* nbits_per_idx (8)                                                */
  int metric;      /* index default metric, GB200_METRIC_*                             */
  int nprobe;      /* index default nprobe (IVFPQModelParams::nprobe, default 80)      */
  int store_raw;   /* 1: keep raw vectors on device (needed for has_rank and FLAT)     */
} gb200_ivfpq_params;

/* Per-call search parameters: IVFPQRetrievalParameters / FlatRetrievalParameters
 * (gamma_index_ivfpq.h:629-673, gamma_index_flat.h:38-64) + the fields of
 * GammaSearchCondition the models read (common/gamma_common_data.h:84-93).           */
typedef struct gb200_search_params {
  int metric;        /* GB200_METRIC_* (retrieval_params metric_type)                 */
  int nprobe;        /* <=0 or >nlist: index default (gamma_index_ivfpq.cc:539-545)   */
  int recall_num;    /* raised to k if smaller (gamma_index_ivfpq.cc:762-765)         */
  int has_rank;      /* 1: exact re-rank of the recall set (compute_dis :646-680)     */
  float min_score;   /* IsSimilarScoreValid window (gamma_common_data.h:95-97)        */
  float max_score;
} gb200_search_params;

/* One range-filter bitmap: RangeQueryResult (table/range_query_result.h:24-160).
 * bit (doc - min_aligned) of `bitmap` set  <=>  doc passes; docs outside [min,max]
 * fail (or pass when not_in).  A search passes all filters (MultiRangeQueryResults::Has,
 * :169-179).  n_filters == 0 means "no range filter" (range_query_result == nullptr).   */
typedef struct gb200_range_filter {
  int min_doc, max_doc; /* min_, max_                                                  */
  int min_aligned;      /* (min_/8)*8                                                  */
  int not_in;           /* b_not_in_                                                   */
  const uint8_t *bitmap;/* ((max_aligned-min_aligned+1)/8) bytes, util/bitmap.cc:25-27 */
} gb200_range_filter;

const char *gb200_last_error(void);
int gb200_device_count(void);

/* ---- lifecycle: RetrievalModel ctor/Init/dtor (retrieval_model.h:220-233) ---------- */
int gb200_ivfpq_create(const gb200_ivfpq_params *p, gb200_index **out);
int gb200_flat_create(int device, int raw_d, int metric, gb200_index **out);
int gb200_destroy(gb200_index *ix);

/* ---- trained state: result of GammaIVFPQIndex::Indexing -> faiss::IndexIVFPQ::train
 * (gamma_index_ivfpq.cc:272-354).  coarse: nlist x d f32 (IndexFlatL2::xb);
 * pq: M x 256 x dsub f32 (ProductQuantizer::centroids).                               */
int gb200_ivfpq_set_quantizers(gb200_index *ix, const float *coarse_centroids,
                               const float *pq_centroids);

/* ---- postings: RTInvertIndex::AddKeys / RealTimeMemData::AddKeys
 * (realtime/realtime_invert_index.cc:43-71, realtime_mem_data.cc:264-303).
 * n postings; posting i goes to the END of list list_no[i] (append order is list
 * order, which is the tie-break order of the scan); vids < 2^31; codes n x M bytes.   */
int gb200_ivfpq_append(gb200_index *ix, int64_t n, const int32_t *list_no,
                       const int64_t *vids, const uint8_t *codes);
/* RealTimeMemData::Update (realtime_mem_data.cc:305-327): same list -> overwrite the
 * code in place; other list -> old posting gets kDelIdxMask (dead), new one appended.   */
int gb200_ivfpq_update(gb200_index *ix, int64_t vid, int32_t new_list_no, const uint8_t *code);
/* list_len[l] as the scan sees it (retrieve_idx_pos_, realtime_invert_index.cc:77-81)  */
int gb200_ivfpq_list_sizes(gb200_index *ix, int64_t *sizes /* nlist */);
/* read one list back in the REFERENCE layout (ids with kDelIdxMask bit 63, AoS codes);
 * test hook mirroring RealTimeMemData::RetrieveCodes (realtime_mem_data.h:95-96).      */
int gb200_ivfpq_get_list(gb200_index *ix, int32_t list_no, int64_t *ids, uint8_t *codes);

/* RealTimeMemData::CompactIfNeed / CompactBucket (realtime/realtime_mem_data.cc:354-424) on the device: drop the
 * postings that were moved away (kDelIdxMask) or whose doc is deleted in the bitmap, keep the survivors' order.
 * list_no >= 0: that list only — it is rewritten into a fresh region and swapped in by publication, searches keep
 * running; list_no == -1: every list, into a new tightly packed pool (also returns the regions abandoned by list
 * growth; searches are drained for the swap).  *dropped (may be NULL) = postings removed.                          */
int gb200_ivfpq_compact(gb200_index *ix, int32_t list_no, int64_t *dropped);
/* Replace the whole content of one list by n postings given in the REFERENCE layout (ids[i] as in idx_array_: vid,
 * bit 63 = kDelIdxMask for a posting that was moved away and only keeps its slot; AoS codes).  This is how a host that
 * owns the lists (the RetrievalModel plugin: RTInvertBucketData after Update / CompactBucket,
 * realtime/realtime_mem_data.cc:119-147, 264-327) brings the device copy of a list in line with its own, whatever
 * happened to it: the new content goes into a fresh region and is swapped in by publication, searches keep running. */
int gb200_ivfpq_replace_list(gb200_index *ix, int32_t list_no, int64_t n, const int64_t *ids, const uint8_t *codes);

/* OPQ pre-transform of the model (model parameter "opq", faiss::OPQMatrix: y = A x + b; index/impl/gamma_index_ivfpq.cc:
 * 158-165 creates it, :338-341 trains it, :448-450 / :547-555 apply it to added vectors and to queries).  Once set,
 * queries are transformed on the device before the coarse quantiser and the table build, added vectors before
 * assign / encode; re-rank keeps using the raw query against the raw vectors (:706).  A: d_out x d_in row-major with
 * d_in == d_out == d; b: d_out floats or NULL.                                                                       */
int gb200_ivfpq_set_opq(gb200_index *ix, int d_in, int d_out, const float *A, const float *b);

/* ---- encode on the device: stage 1 of GammaIVFPQIndex::Add (index/impl/gamma_index_ivfpq.cc:424-476) —
 * quantizer->assign (the tensor-core coarse stage with nprobe = 1), compute_residuals, pq.compute_codes with faiss'
 * own sub-distance arithmetic (codes are bit-identical to the CPU engine's for nsubvector slices narrower than 16
 * floats; an assignment can differ only where two centroids are at rounding distance).
 * x: n rows of x_dim floats, x_dim <= d (columns beyond x_dim are zero: ConvertVectorDim).  list_no n, codes n x M. */
int gb200_ivfpq_encode(gb200_index *ix, int64_t n, const float *x, int x_dim, int32_t *list_no, uint8_t *codes);
/* the whole Add for vids first_vid .. first_vid + n - 1: upload the raw rows (n x raw_d), encode them on the device,
 * append the postings (AddKeys semantics as gb200_ivfpq_append).  list_no / codes (may be NULL) return what was
 * appended, for a host that keeps its own copy of the lists (Dump).                                                  */
int gb200_ivfpq_add_raw(gb200_index *ix, int64_t first_vid, int64_t n, const float *x, int32_t *list_no, uint8_t *codes);
/* same for rows that are already in the device raw store (gb200_upload_raw / gb200_upload_raw_dev): encode + append
 * vids first_vid .. first_vid + n - 1 — the engine's AddRTVecsToIndex loop, which indexes what AddToStore stored
 * earlier (vector/vector_manager.cc:280-382).  list_no / codes may be NULL.                                            */
int gb200_ivfpq_add_stored(gb200_index *ix, int64_t first_vid, int64_t n, int32_t *list_no, uint8_t *codes);

/* ---- raw vectors: the read side of VectorReader::Gets / RawVector::GetVectorHeader
 * (index/retrieval_model.h:192-215, vector/memory_raw_vector.cc:110-142); vids are
 * implicit = first_vid .. first_vid+n-1; re-upload of an existing range = UpdateToStore. */
int gb200_upload_raw(gb200_index *ix, int64_t first_vid, int64_t n, const float *x);
/* same, rows already in device memory (a loader that decodes on the GPU, bench tooling)  */
int gb200_upload_raw_dev(gb200_index *ix, int64_t first_vid, int64_t n, const float *x_dev);
int64_t gb200_raw_count(gb200_index *ix);

/* ---- deleted-docs bitmap: bitmap::BitmapManager::Set/Unset (util/bitmap_manager.cc),
 * bit = 1 => deleted; consulted by GammaSearchCondition::IsValid (:99-108).            */
int gb200_set_deleted(gb200_index *ix, const int64_t *docids, int64_t n, int deleted);
/* bring the device copy in line with the whole reference bitmap (bit = 1: deleted).  Only words that differ from the
 * library's shadow copy travel, so a model may call this before every Search: the reference tests the bitmap live and
 * some engine paths set bits without RetrievalModel::Delete (search/gamma_engine.cc:866).                           */
int gb200_upload_deleted_bitmap(gb200_index *ix, const uint8_t *bitmap, int64_t nbits);

/* ---- search: RetrievalModel::Search (retrieval_model.h:282-284).
 * THREADING: every entry point may be called concurrently from any number of threads (the engine's request threads
 * all call Search on one model, gamma_engine.cc:74-97, tests/test.h:1033-1062).  Each Search call runs on its own
 * stream and workspaces (up to GB200_MAX_CONTEXTS calls in flight per index, further callers wait for a free one);
 * appends / updates / deletes / raw uploads are serialised among themselves and do not block searches except for
 * the moment a device array has to be re-allocated.  A search sees a list either before or after an append.
 * Concurrent gb200_ivfpq_search calls with equal parameters and no range filter may travel in one device batch
 * (GB200_COALESCE); each caller's rows are bit-identical to a call of its own.
 * GammaIVFPQIndex::Search (gamma_index_ivfpq.cc:514-566): coarse quantizer, ADC scan of
 * the nprobe lists with the validity filter inside the scan, recall_num selection,
 * optional exact re-rank, score window, top-k.  xq: n x d f32; out: n x k.             */
int gb200_ivfpq_search(gb200_index *ix, int n, const float *xq, int k,
                       const gb200_search_params *sp, const gb200_range_filter *filters,
                       int n_filters, float *distances, int64_t *labels);
/* GammaIVFPQIndex::search_preassigned (gamma_index_ivfpq.cc:701-890): probes given.
 * keys n x nprobe (i64, -1 = none), coarse_dis n x nprobe.                             */
int gb200_ivfpq_search_preassigned(gb200_index *ix, int n, const float *xq, int k,
                                   const gb200_search_params *sp,
                                   const gb200_range_filter *filters, int n_filters,
                                   const int64_t *keys, const float *coarse_dis, int nprobe,
                                   float *distances, int64_t *labels);
/* coarse stage alone: quantizer->search (gamma_index_ivfpq.cc:560).                    */
int gb200_ivfpq_coarse(gb200_index *ix, int n, const float *xq, int nprobe,
                       float *coarse_dis, int64_t *keys);
/* GammaFLATIndex::Search (gamma_index_flat.cc:118-300) over the uploaded raw vectors;
 * also the brute_force_search / untrained fallback of the IVFPQ model (:529-537).      */
int gb200_flat_search(gb200_index *ix, int n, const float *xq, int k,
                      const gb200_search_params *sp, const gb200_range_filter *filters,
                      int n_filters, float *distances, int64_t *labels);

/* ---- device-resident variants (queries/results already in HBM; stream = cudaStream_t
 * as void*).  Used by bench.py's device-timed leg and by the multi-GPU driver so the
 * per-rank top-k can feed ncclAllGather without a host round trip.                     */
int gb200_ivfpq_search_dev(gb200_index *ix, int n, const float *xq_dev, int k,
                           const gb200_search_params *sp, float *distances_dev,
                           int64_t *labels_dev, void *stream);
int gb200_flat_search_dev(gb200_index *ix, int n, const float *xq_dev, int k,
                          const gb200_search_params *sp, float *distances_dev,
                          int64_t *labels_dev, void *stream);
/* install a per-index "valid docs" filter for the *_dev calls (same semantics as the
 * filters argument above); n_filters = 0 clears it.                                    */
int gb200_set_filters(gb200_index *ix, const gb200_range_filter *filters, int n_filters);

/* ---- accounting: RetrievalModel::GetTotalMemBytes (retrieval_model.h:287) + bench ---- */
int64_t gb200_mem_bytes(gb200_index *ix);
/* postings the last IVFPQ search scanned (sum over (query,probe) of list length) and the
 * number of kernels it launched — bench.py's algorithmic-bytes and gpu_launches.        */
int64_t gb200_last_scanned_postings(gb200_index *ix);
int64_t gb200_launch_count(gb200_index *ix);
/* device time (ms, CUDA events on the search stream) of the stages of the last search:
 * out[0]=coarse, out[1]=scan, out[2]=merge/rerank, out[3]=total.                        */
int gb200_last_stage_ms(gb200_index *ix, float *out4);
int gb200_set_profiling(gb200_index *ix, int enable);
/* device time (ms) of the ADC scan kernel alone in the last IVFPQ search (CUDA events around that one launch on the
 * search stream; profiling must be enabled) — the duration bench.py divides the algorithmic bytes by.            */
float gb200_last_scan_kernel_ms(gb200_index *ix);
/* wait for the index's stream (after *_dev calls) and refresh the counters above.          */
int gb200_sync(gb200_index *ix);
/* Tuning knobs (GB200_* environment variables) are read once when an index is created; this re-reads
 * them for an existing index (A/B runs in bench.py and the tests).                           */
int gb200_reload_tuning(gb200_index *ix);

/* ---- IVFFLAT: the reference's "IVFFLAT" model (index/impl/gamma_index_ivfflat.{h,cc}: GammaIndexIVFFlat — IVF over raw
 * float lists, exact distances, no re-rank).  The device keeps one copy of every vector (the raw store, by vid) and
 * the inverted lists hold vids, so Add = gb200_upload_raw + an assignment.  Realtime list maintenance, deleted
 * bitmap, filters and the multi-GPU exchange are the IVFPQ model's (gb200_ivfpq_update is not meaningful here; use
 * gb200_ivfpq_replace_list with any code bytes, gb200_ivfpq_compact, gb200_set_deleted, ...).                        */
int gb200_ivfflat_create(int device, int d, int nlist, int metric, int nprobe, gb200_index **out);
/* coarse centroids (IndexFlatL2 xb after train), nlist x d                                                           */
int gb200_ivfflat_set_quantizer(gb200_index *ix, const float *coarse);
/* RTInvertIndex::AddKeys for vectors already in the raw store: list_no n, vids n                                     */
int gb200_ivfflat_append(gb200_index *ix, int64_t n, const int32_t *list_no, const int64_t *vids);
/* GammaIndexIVFFlat::Add (gamma_index_ivfflat.cc:265-330): upload n x d rows as vids first_vid.., assign them with the
 * tensor-core coarse stage (quantizer->assign) and append; list_no (may be NULL) returns the assignment               */
int gb200_ivfflat_add_raw(gb200_index *ix, int64_t first_vid, int64_t n, const float *x, int32_t *list_no);
/* GammaIndexIVFFlat::Search (:392-421) + search_preassigned (:423-560); sp->recall_num / has_rank are ignored          */
int gb200_ivfflat_search(gb200_index *ix, int n, const float *xq, int k, const gb200_search_params *sp,
                         const gb200_range_filter *filters, int n_filters, float *distances, int64_t *labels);

/* ---- multi-GPU: one process per GPU, index replicated, batch sharded by query.  Replaces the host-side
 * merge of faiss IndexReplicas / IndexShards the reference's GPU model relies on (index/impl/gpu/gamma_gpu_cloner.cpp:
 * 209-212).  Every rank owns a result window in device memory; after its search it stores its [n][k] block into every
 * peer's window over NVLink (peer memory mapped through CUDA IPC) and waits for theirs — one kernel, no collective
 * library on the data path.  The opaque handles travel between the processes by whatever channel the host has
 * (MPI, torch.distributed, a file).                                                                                  */
typedef struct gb200_comm gb200_comm;
#define GB200_COMM_HANDLE_BYTES 128
/* slot_bytes >= n * k * 12 of the largest per-rank batch; handle: GB200_COMM_HANDLE_BYTES bytes to hand to the peers */
int gb200_comm_create(int device, int rank, int world, int64_t slot_bytes, gb200_comm **out, uint8_t *handle);
/* handles: world x GB200_COMM_HANDLE_BYTES, rank-major (this rank's own entry is ignored) */
int gb200_comm_connect(gb200_comm *c, const uint8_t *handles);
int gb200_comm_destroy(gb200_comm *c);
int64_t gb200_comm_slot_bytes(gb200_comm *c);
/* 0 = every exchange so far saw all peers; 1 + p = peer p did not arrive within ~4 s (synchronises the device)       */
int gb200_comm_status(gb200_comm *c);
/* device pointers for the NEXT exchange: where this rank's result goes ([n*k] f32 then [n*k] i64) and the start of the
 * gathered window (rank r's block at all_slots + r * slot_bytes)                                                       */
int gb200_comm_buffers(gb200_comm *c, void **my_slot, void **all_slots);
/* copy (part of) the gathered window to host memory on `stream`; sync != 0 also waits for the copy                     */
int gb200_comm_read(gb200_comm *c, void *dst_host, const void *src_dev, int64_t bytes, void *stream, int sync);
/* push the block written into my_slot to every peer and wait for theirs, all on `stream`                              */
int gb200_comm_exchange(gb200_comm *c, int64_t bytes, void *stream);
/* search this rank's n queries (device pointers) and exchange: afterwards (in stream order) rank r's distances are at
 * *D_all + r * slot_bytes / 4 ... — see gb200_comm_buffers for the layout                                             */
int gb200_ivfpq_search_sharded(gb200_index *ix, gb200_comm *c, int n, const float *xq_dev, int k,
                               const gb200_search_params *sp, float **D_all, int64_t **I_all_of_rank0, void *stream);
/* the pipelined form: search, push this rank's result to the peers, and wait only for the peers' results of the PREVIOUS
 * call — *D_all_prev is that call's gathered window (NULL on the first call).  No rank idles for the slowest rank of the
 * current step; a server consumes gathered results one call late.  gb200_comm_flush waits (on `stream`) for the last
 * call's results and returns their window.  Both forms may be mixed on one gb200_comm.                                */
int gb200_ivfpq_search_sharded_deferred(gb200_index *ix, gb200_comm *c, int n, const float *xq_dev, int k,
                                        const gb200_search_params *sp, float **D_all_prev, int64_t **I_all_prev_of_rank0,
                                        void *stream);
int gb200_comm_flush(gb200_comm *c, void **all_slots, void *stream);

/* ---- test hook (not used by the plugin): run the streaming top-R selection primitive the scan
 * kernels use (append + radix-select prune, one CTA of `threads`) on caller-provided 64-bit keys fed
 * in batches, return the R smallest (unordered).  Lets tests check the k-select against a host
 * sort in isolation — the role faiss' own Heap tests play for heap_replace_top.                  */
int gb200_debug_select(int device, const uint64_t *keys, int n, int R, int cap, int batch, int threads,
                       uint64_t *out, int *out_n);

/* test hook (pure host arithmetic, no device needed): the positional work plan of one IVFPQ scan launch for a batch of
 * n queries on `slots` resident CTAs — out4 = {n_full, s_tail, n_items, rows}: queries [0, n_full) are one work item
 * each, the rest s_tail items each; rows = candidate rows per query handed to the re-rank.            */
int gb200_debug_plan(int n, int slots, int nprobe, int recall_num, int s_uniform, int *out4);

#ifdef __cplusplus
}
#endif
#endif /* GAMMA_B200_H_ */
