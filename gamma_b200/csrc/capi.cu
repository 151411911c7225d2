// extern "C" shim of include/gamma_b200.h: owns the device mirror of one RetrievalModel
// (quantizers, realtime posting pools, raw vectors, live-docs bitmap) and sequences the
// kernels of a Search call.  No CPU fallback exists anywhere in here: every data-path step
// is a kernel launch; the host only moves bytes and keeps list extents.
//
// Threading (reference: Search is called concurrently from many request threads while one thread
// Adds / Updates and others Delete — SURVEY §8b, tests/test.h:1033-1062, search/gamma_engine.cc:74-97):
//   * every Search call takes a SearchCtx (own stream, side stream, events, all workspaces) from a
//     small pool, so concurrent callers overlap on the device instead of queueing behind one mutex;
//   * searches hold data_mu SHARED.  Writers (append / update / raw upload / delete) are serialised
//     among themselves by writer_mu and also hold data_mu shared: they write behind the published
//     list lengths on their own stream and publish extents with release semantics
//     (postings.cu publish_lists_kernel), so a search never waits for a writer;
//   * only operations that free or replace device arrays (pool / raw store / bitmap growth, set_quantizers,
//     full compaction, destroy) take data_mu EXCLUSIVE, after draining the device.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <thread>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "../../include/gamma_b200.h"
#include "coalesce.h"
#include "kernels.h"

using namespace gb;

static thread_local std::string g_err;
static void set_err(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}
#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      set_err("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
      return e_ == cudaErrorMemoryAllocation ? GB200_ENOMEM : GB200_ECUDA;               \
    }                                                                                    \
  } while (0)
#define CKI(expr)                  \
  do {                             \
    int r_ = (expr);               \
    if (r_ != GB200_OK) return r_; \
  } while (0)

namespace {

struct DevBuf {  // grow-only device scratch
  void *p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return GB200_OK;
    if (p) cudaFree(p);  // synchronises with the device: nothing can still be reading the old buffer
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      set_err("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return GB200_ENOMEM;
    }
    cap = want;
    return GB200_OK;
  }
  template <typename T>
  T *as() { return reinterpret_cast<T *>(p); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

inline long long roundup32(long long v) { return (v + 31) & ~31LL; }

// Tuning knobs: environment variables read ONCE at index creation (gb200_reload_tuning re-reads them, for A/B runs);
// defaults are the measured best.
struct Tuning {
  int scan_threads = 0;   // GB200_SCAN_THREADS: M = 32 kernel: 384 (default) / 320 / 256 (2 CTAs / SM, ring 2 / 3 / 4), 512
  int pf_blocks = 4;      // GB200_SCAN_PF: M = 64 kernel: L2 prefetch distance in 32-posting blocks
  int ch_blocks = 8;      // GB200_SCAN_CH: blocks per item (v3)
  int help_min = 8;       // GB200_SCAN_HELP_MIN: idle CTAs join a query that has >= this many unclaimed items (v3)
  int max_rows = 0;       // GB200_SCAN_ROWS: candidate rows per query (v3), 0 = automatic
  int v3_flags = 4;       // GB200_SCAN_FLAGS: M = 32 scan, bit 0: next query fetched inside the scan, bit 1: its tables
                          // requested before the final select, bit 2: approximate in-loop prunes
  int rerank_no_stage = 1;  // GB200_RERANK_STAGE=1: re-rank stages the raw rows in shared memory by cp.async (validated, slower:
                            // 70 us against 42 — three CTAs per SM instead of sixteen) instead of L2 prefetch + load on use
  int scan_cap = 0;       // GB200_SCAN_CAP: candidate buffer of the M = 32 scan in keys (0 = automatic)
  int v3_tma = 0;         // GB200_SCAN_TMA: v3 posting ring fed by bulk copies (1) or per-lane cp.async (0)
  int splits = 0;         // GB200_SCAN_SPLITS: M = 64 / generic: CTAs per query, 0 = automatic
  int tail = 0;           // GB200_SCAN_TAIL: M = 64 plan: splits of the last partial wave, 0 = automatic
  int no_plan = 0;        // GB200_SCAN_NOPLAN: M = 64 without the positional work plan
  int coarse_simt = 0;    // GB200_COARSE=simt: CUDA-core fp32 coarse distances instead of the tcgen05 GEMM
  int coarse_full_select = 0;  // GB200_COARSE_FULL_SELECT=1: select over whole rows instead of starting from chunk minima
  int lut_inline = 0;     // GB200_LUT_INLINE: build the tables on the main stream
  int flat_mode = 0;      // GB200_FLAT: 0 automatic, 1 = exact (per-query scan), 2 = tc (tensor-core path for any batch)
  int max_contexts = 8;   // GB200_MAX_CONTEXTS: Search calls in flight per index (each owns streams + workspaces)
  int coalesce = 3;       // GB200_COALESCE: concurrent host-buffer IVFPQ searches with equal parameters are merged into one
                          // device batch once this many merged batches are already in flight (0 = never merge)
  int coalesce_max = 2048;  // GB200_COALESCE_MAX: queries per merged batch (calls of at least half this size are not merged)
  int coalesce_wait_us = 60;  // GB200_COALESCE_WAIT_US: while another batch keeps the device busy, a starting batch waits up
                              // to this long for the callers it expects (recent concurrency / coalesce); 0 = never waits
  int coalesce_balance = 1;   // GB200_COALESCE_BALANCE: a batch takes at most its share (recent concurrency / coalesce)
  long long flat_chunk_rows = 0;  // GB200_FLAT_CHUNK_ROWS: database rows per tensor-core chunk (tests: force many chunks)
  void read() {
    *this = Tuning();
    auto geti = [](const char *k, int d) { const char *e = getenv(k); return e && *e ? atoi(e) : d; };
    scan_threads = geti("GB200_SCAN_THREADS", scan_threads);
    if (scan_threads != 256 && scan_threads != 320 && scan_threads != 384 && scan_threads != 416 && scan_threads != 448 &&
        scan_threads != 512)
      scan_threads = 0;
    pf_blocks = std::max(0, geti("GB200_SCAN_PF", pf_blocks));
    ch_blocks = std::max(1, geti("GB200_SCAN_CH", ch_blocks));
    help_min = std::max(1, geti("GB200_SCAN_HELP_MIN", help_min));
    v3_tma = geti("GB200_SCAN_TMA", v3_tma);
    scan_cap = std::max(0, geti("GB200_SCAN_CAP", scan_cap));
    v3_flags = geti("GB200_SCAN_FLAGS", v3_flags);
    rerank_no_stage = geti("GB200_RERANK_STAGE", 0) ? 0 : 1;
    max_rows = std::max(0, geti("GB200_SCAN_ROWS", 0));
    splits = std::max(0, geti("GB200_SCAN_SPLITS", 0));
    tail = std::max(0, geti("GB200_SCAN_TAIL", 0));
    no_plan = geti("GB200_SCAN_NOPLAN", 0);
    lut_inline = geti("GB200_LUT_INLINE", 0);
    coarse_full_select = geti("GB200_COARSE_FULL_SELECT", 0);
    max_contexts = std::max(1, std::min(64, geti("GB200_MAX_CONTEXTS", max_contexts)));
    coalesce = std::max(0, std::min(max_contexts, geti("GB200_COALESCE", coalesce)));
    coalesce_max = std::max(2, geti("GB200_COALESCE_MAX", coalesce_max));
    coalesce_wait_us = std::max(0, std::min(1000, geti("GB200_COALESCE_WAIT_US", coalesce_wait_us)));
    coalesce_balance = geti("GB200_COALESCE_BALANCE", coalesce_balance);
    if (const char *e = getenv("GB200_COARSE")) coarse_simt = !strcmp(e, "simt");
    if (const char *e = getenv("GB200_FLAT")) flat_mode = !strcmp(e, "exact") ? 1 : !strcmp(e, "tc") ? 2 : 0;
    if (const char *e = getenv("GB200_FLAT_CHUNK_ROWS")) flat_chunk_rows = atoll(e);
  }
};

// Everything one in-flight Search call owns.  All uses of its buffers are ordered on `stream` (the side stream joins
// through events), so a context can be handed to the next call while its last kernels still run.
struct SearchCtx {
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // side stream: per-query lookup tables are built while the coarse quantiser runs
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_user = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // 4, 5 bracket the scan kernel alone
  DevBuf ws_xq, ws_xn, ws_dist, ws_keys, ws_cdis, ws_cand, ws_out_d, ws_out_i, ws_flat, ws_lut, ws_xs, ws_fstate, ws_probe,
      ws_ctl, ws_cmin, ws_xt;
  DevBuf valid_filt, filt_bytes, filt_desc;  // per-call range filters -> validity bitmap
  unsigned long long *d_scanned = nullptr;
  void *h_stage = nullptr;  // pinned: results of a merged batch before they are handed to their callers
  size_t h_stage_cap = 0;
  int lut_built_n = 0, lut_built_ip = -1;  // the side stream holds tables for this many queries of the current search
  int zero_req = 0, zeroed_for = 0;  // scan control words: queries the coarse stage should zero them for / did
  const gb::PeerSink *sink = nullptr;  // multi-GPU: the final kernel of this search also feeds the peers (comm.cu)
  bool sink_used = false;
  long long launches = 0;
  bool timed = false;  // the events of the last call were recorded

  int init() {
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_user, cudaEventDisableTiming));
    for (int i = 0; i < 6; i++) CK(cudaEventCreate(&ev[i]));
    CK(cudaMalloc(&d_scanned, sizeof(unsigned long long)));
    CK(cudaMemset(d_scanned, 0, sizeof(unsigned long long)));
    return GB200_OK;
  }
  void destroy() {
    if (stream) cudaStreamSynchronize(stream);
    DevBuf *bufs[] = {&ws_xq,  &ws_xn, &ws_dist,   &ws_keys,  &ws_cdis, &ws_cand,    &ws_out_d,   &ws_out_i, &ws_flat,
                      &ws_lut, &ws_xs, &ws_fstate, &ws_probe, &ws_ctl,  &valid_filt, &filt_bytes, &filt_desc, &ws_cmin, &ws_xt};
    for (DevBuf *b : bufs) b->release();
    if (d_scanned) cudaFree(d_scanned);
    if (h_stage) cudaFreeHost(h_stage);
    for (int i = 0; i < 6; i++)
      if (ev[i]) cudaEventDestroy(ev[i]);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_user) cudaEventDestroy(ev_user);
    if (stream2) cudaStreamDestroy(stream2);
    if (stream) cudaStreamDestroy(stream);
  }
};

}  // namespace

struct gb200_index {
  int kind = 0;  // 0 = IVFPQ, 1 = FLAT
  Tuning tune;
  gb200_ivfpq_params p{};
  int dsub = 0, chunk = 0, layout = 0, mode = 0;
  bool flat_lists = false;  // IVFFLAT: the lists carry vids only (4 dummy code bytes), distances come from the raw store
  int smem_reserved = 0;  // cudaDevAttrReservedSharedMemoryPerBlock (the M = 32 / 64 kernels' LDS immediates assume 1024)
  int num_sms = 0;        // cudaDeviceProp::multiProcessorCount (grid sizing of the persistent scan, work plans)

  // ---- locking (see the header comment)
  std::shared_mutex data_mu;
  std::mutex writer_mu;
  std::mutex ctx_mu;
  std::condition_variable ctx_cv;
  // concurrent small searches waiting to be merged into one device batch (search_coalesced, coalesce.h)
  struct PendingSearch {
    int n, k;
    const float *xq;
    float *D;
    int64_t *I;
    gb200_search_params sp;
    bool same(const PendingSearch &b) const {
      return k == b.k && sp.metric == b.sp.metric && sp.nprobe == b.sp.nprobe && sp.recall_num == b.sp.recall_num &&
             sp.has_rank == b.sp.has_rank && sp.min_score == b.sp.min_score && sp.max_score == b.sp.max_score;
    }
  };
  gb::Coalescer<PendingSearch> coalescer;
  std::vector<SearchCtx *> ctx_all, ctx_free;
  SearchCtx *last_ctx = nullptr;  // most recently used context (gb200_sync / profiling read-out)
  std::mutex stats_mu;

  // ---- writer side: own stream + staging
  cudaStream_t wstream = nullptr;
  DevBuf w_stage, w_pub;
  long long *d_woff = nullptr;  // the writer's view of the list offsets (regions being filled before they are published)

  std::atomic<bool> trained{false};
  float *d_cent = nullptr, *d_cent_norm = nullptr, *d_pq = nullptr, *d_pq_t = nullptr;
  float *d_cent_small = nullptr;  // centroid - tf32(centroid): second operand of the 3xTF32 tensor-core GEMM
  // OPQ pre-transform (gb200_ivfpq_set_opq): A transposed [d][d] and the bias, or nullptr
  float *d_opq_At = nullptr, *d_opq_b = nullptr;

  // posting pools: bump allocation, lists grow by allocate-copy-swap at the tail (host tables are writer-only)
  long long pool_cap = 0, pool_used = 0, pool_live_cap = 0;
  uint8_t *d_codes = nullptr;
  int *d_ids = nullptr;
  float *d_norms = nullptr;
  std::vector<long long> h_off;
  std::vector<int> h_len, h_cap;
  long long *d_off = nullptr;
  int *d_len = nullptr;
  std::vector<long long> vid_loc;
  std::atomic<long long> max_vid{-1};
  std::atomic<int> max_list_len{0};  // longest list ever published (upper bound; sizes the scan's per-query item table)

  float *d_raw = nullptr;
  long long raw_cap = 0;
  std::atomic<long long> raw_n{0};
  // lazily built companions of the raw store for the tensor-core flat path: x - tf32(x) and |x|^2
  float *d_raw_small = nullptr, *d_raw_norm = nullptr;
  long long aux_cap = 0;
  std::atomic<long long> aux_n{0};
  std::mutex aux_mu;

  // live-docs bitmap: bit = 1 <=> NOT deleted (the complement of bitmap::BitmapManager's bits); words beyond the
  // highest doc ever touched are all ones.  h_deleted is the writer's shadow of the reference bitmap.  While any doc is
  // deleted the bitmap covers at least doc_bits() docs (writers grow it BEFORE they publish new docs).
  std::vector<uint32_t> h_deleted;
  uint32_t *d_live = nullptr;
  long long live_words = 0;  // capacity of d_live in words
  std::atomic<long long> deleted_count{0};
  // range filters installed for the *_dev calls (gb200_set_filters): host copy + the bitmap built from it
  struct Installed {
    bool active = false;
    std::vector<gb200_range_filter> desc;
    std::vector<std::vector<uint8_t>> bytes;
    DevBuf valid, filt_bytes, filt_desc;
    long long bits = 0;
  } inst;

  // last-call statistics
  bool profiling = false;
  long long last_scanned = 0;
  std::atomic<long long> launches{0};
  float stage_ms[4] = {0, 0, 0, 0};
  float scan_kernel_ms = 0;

  long long doc_bits() const { return std::max(max_vid.load() + 1, raw_n.load()); }
};

static int use_device(gb200_index *ix) {
  CK(cudaSetDevice(ix->p.device));
  return GB200_OK;
}

// ---- search contexts ------------------------------------------------------------------------------------
static SearchCtx *acquire_ctx(gb200_index *ix) {
  std::unique_lock<std::mutex> g(ix->ctx_mu);
  for (;;) {
    if (!ix->ctx_free.empty()) {
      SearchCtx *c = ix->ctx_free.back();
      ix->ctx_free.pop_back();
      return c;
    }
    if ((int)ix->ctx_all.size() < ix->tune.max_contexts) {
      SearchCtx *c = new SearchCtx;
      if (c->init() != GB200_OK) {
        c->destroy();
        delete c;
        return nullptr;
      }
      ix->ctx_all.push_back(c);
      return c;
    }
    ix->ctx_cv.wait(g);
  }
}
static void release_ctx(gb200_index *ix, SearchCtx *c) {
  {
    std::lock_guard<std::mutex> g(ix->ctx_mu);
    ix->ctx_free.push_back(c);
    ix->last_ctx = c;
  }
  ix->ctx_cv.notify_one();
}
namespace {
struct SearchScope {  // shared hold on the index data + one context, for the duration of a Search entry point
  gb200_index *ix;
  std::shared_lock<std::shared_mutex> lk;
  SearchCtx *c;
  explicit SearchScope(gb200_index *ix_) : ix(ix_), lk(ix_->data_mu), c(acquire_ctx(ix_)) {}
  ~SearchScope() {
    if (c) release_ctx(ix, c);
  }
};
// exclusive hold: every search has left its entry point; kernels they enqueued (the *_dev calls return before their
// work is done) are drained before anything is freed or replaced
struct ExclusiveScope {
  std::unique_lock<std::shared_mutex> lk;
  explicit ExclusiveScope(gb200_index *ix) : lk(ix->data_mu) { cudaDeviceSynchronize(); }
};
}  // namespace

extern "C" {

const char *gb200_last_error(void) { return g_err.c_str(); }

int gb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

static int common_create(gb200_index *ix) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ix->p.device < 0 || ix->p.device >= ndev) {
    set_err("device %d not present (%d devices)", ix->p.device, ndev);
    return GB200_EINVAL;
  }
  CK(cudaSetDevice(ix->p.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ix->p.device));
  if (prop.major < 10) {
    set_err("device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", ix->p.device, prop.major,
            prop.minor);
    return GB200_EUNSUPPORTED;
  }
  ix->smem_reserved = (int)prop.reservedSharedMemPerBlock;
  ix->num_sms = prop.multiProcessorCount;
  ix->tune.read();
  CK(cudaStreamCreateWithFlags(&ix->wstream, cudaStreamNonBlocking));
  return GB200_OK;
}

static void free_index(gb200_index *ix) {
  cudaSetDevice(ix->p.device);
  cudaDeviceSynchronize();
  for (SearchCtx *c : ix->ctx_all) {
    c->destroy();
    delete c;
  }
  void *ptrs[] = {ix->d_raw_small, ix->d_raw_norm, ix->d_cent_small, ix->d_cent, ix->d_cent_norm, ix->d_pq,   ix->d_pq_t, ix->d_codes,
                  ix->d_ids,       ix->d_norms,    ix->d_off,        ix->d_len,  ix->d_woff,      ix->d_raw,  ix->d_live,
                  ix->d_opq_At,    ix->d_opq_b};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  DevBuf *bufs[] = {&ix->w_stage, &ix->w_pub, &ix->inst.valid, &ix->inst.filt_bytes, &ix->inst.filt_desc};
  for (DevBuf *b : bufs) b->release();
  if (ix->wstream) cudaStreamDestroy(ix->wstream);
  cudaGetLastError();
  delete ix;
}

int gb200_ivfpq_create(const gb200_ivfpq_params *p, gb200_index **out) {
  if (!p || !out) return GB200_EINVAL;
  if (p->d <= 0 || p->nlist <= 0 || p->nsubvector <= 0 || p->raw_d <= 0 || p->raw_d > p->d) {
    set_err("bad ivfpq params");
    return GB200_EINVAL;
  }
  if (p->nbits != 8) {
    set_err("nbits_per_idx=%d: only 8 is implemented", p->nbits);
    return GB200_EUNSUPPORTED;
  }
  if (p->d % p->nsubvector != 0 || p->nsubvector % 4 != 0) {
    set_err("need d %% nsubvector == 0 and nsubvector %% 4 == 0 (d=%d M=%d)", p->d, p->nsubvector);
    return GB200_EUNSUPPORTED;
  }
  gb200_index *ix = new gb200_index;
  ix->kind = 0;
  ix->p = *p;
  ix->dsub = p->d / p->nsubvector;
  int M = p->nsubvector;
  ix->chunk = (M % 16 == 0) ? 16 : (M % 8 == 0 ? 8 : 4);
  int rc = common_create(ix);
  if (rc != GB200_OK) {
    free_index(ix);
    return rc;
  }
  // posting layout / scan kernel, fixed at creation: M = 32 and M = 64 have conflict-free kernels over a pre-rotated
  // layout (LAYOUT_M32_ROT / LAYOUT_M64_ROT); every other M % 4 == 0 runs the generic kernel over the plain layout
  // (GB200_FORCE_GENERIC=1 forces that one, for A/B and parity tests).
  const char *force = getenv("GB200_FORCE_GENERIC");
  const bool generic = (force && force[0] == '1') || p->d > 1024 || ix->smem_reserved != 1024;  // the kernels' LDS immediates
  ix->layout = LAYOUT_PLAIN;
  ix->mode = 0;
  if (M == 32 && !generic) ix->layout = LAYOUT_M32_ROT, ix->mode = 1;
  if (M == 64 && !generic) ix->layout = LAYOUT_M64_ROT, ix->mode = 2;
  ix->h_off.assign(p->nlist, 0);
  ix->h_len.assign(p->nlist, 0);
  ix->h_cap.assign(p->nlist, 0);
  if (cudaMalloc(&ix->d_off, sizeof(long long) * p->nlist) != cudaSuccess ||
      cudaMalloc(&ix->d_woff, sizeof(long long) * p->nlist) != cudaSuccess ||
      cudaMalloc(&ix->d_len, sizeof(int) * p->nlist) != cudaSuccess) {
    set_err("alloc list tables");
    free_index(ix);
    return GB200_ENOMEM;
  }
  cudaMemset(ix->d_off, 0, sizeof(long long) * p->nlist);
  cudaMemset(ix->d_woff, 0, sizeof(long long) * p->nlist);
  cudaMemset(ix->d_len, 0, sizeof(int) * p->nlist);
  *out = ix;
  return GB200_OK;
}

int gb200_flat_create(int device, int raw_d, int metric, gb200_index **out) {
  if (!out || raw_d <= 0) return GB200_EINVAL;
  gb200_index *ix = new gb200_index;
  ix->kind = 1;
  ix->p.device = device;
  ix->p.d = raw_d;
  ix->p.raw_d = raw_d;
  ix->p.metric = metric;
  ix->p.store_raw = 1;
  int rc = common_create(ix);
  if (rc != GB200_OK) {
    free_index(ix);
    return rc;
  }
  *out = ix;
  return GB200_OK;
}

int gb200_destroy(gb200_index *ix) {
  if (!ix) return GB200_OK;
  {
    // searches and writers that are inside the library drain first; a caller that enters after this point is using a
    // destroyed handle (its bug, as with any destructor)
    std::lock_guard<std::mutex> w(ix->writer_mu);
    cudaSetDevice(ix->p.device);
    ExclusiveScope x(ix);
  }
  free_index(ix);
  return GB200_OK;
}

int gb200_ivfpq_set_quantizers(gb200_index *ix, const float *coarse, const float *pq) {
  if (!ix || ix->kind != 0 || !coarse || !pq) return GB200_EINVAL;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  ExclusiveScope x(ix);
  const int d = ix->p.d, nlist = ix->p.nlist, M = ix->p.nsubvector, dsub = ix->dsub;
  size_t cb = (size_t)nlist * d * sizeof(float), pb = (size_t)M * 256 * dsub * sizeof(float);
  if (!ix->d_cent) {
    CK(cudaMalloc(&ix->d_cent, cb));
    CK(cudaMalloc(&ix->d_cent_norm, (size_t)nlist * sizeof(float)));
    CK(cudaMalloc(&ix->d_pq, pb));
    CK(cudaMalloc(&ix->d_pq_t, pb));
  }
  CK(cudaMemcpyAsync(ix->d_cent, coarse, cb, cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaMemcpyAsync(ix->d_pq, pq, pb, cudaMemcpyHostToDevice, ix->wstream));
  // code-major copy [256][M][dsub] for the table build
  std::vector<float> t((size_t)M * 256 * dsub);
  for (int m = 0; m < M; m++)
    for (int c = 0; c < 256; c++)
      memcpy(&t[((size_t)c * M + m) * dsub], &pq[((size_t)m * 256 + c) * dsub], dsub * sizeof(float));
  CK(cudaMemcpyAsync(ix->d_pq_t, t.data(), pb, cudaMemcpyHostToDevice, ix->wstream));
  CK(launch_row_norms(ix->d_cent, nlist, d, ix->d_cent_norm, ix->wstream));
  if (!ix->d_cent_small) CK(cudaMalloc(&ix->d_cent_small, cb));
  CK(launch_tf32_residual(ix->d_cent, ix->d_cent_small, (size_t)nlist * d, ix->wstream));
  ix->launches += 2;
  CK(cudaStreamSynchronize(ix->wstream));
  ix->trained = true;
  return GB200_OK;
}

int gb200_ivfpq_set_opq(gb200_index *ix, int d_in, int d_out, const float *A, const float *b) {
  if (!ix || ix->kind != 0 || !A) return GB200_EINVAL;
  if (d_in != ix->p.d || d_out != ix->p.d) {
    set_err("opq %d -> %d on an index of dimension %d: only d_in == d_out == d is implemented", d_in, d_out, ix->p.d);
    return GB200_EUNSUPPORTED;
  }
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  ExclusiveScope x(ix);
  const int d = ix->p.d;
  std::vector<float> At((size_t)d * d);
  for (int o = 0; o < d; o++)
    for (int k = 0; k < d; k++) At[(size_t)k * d + o] = A[(size_t)o * d + k];
  if (!ix->d_opq_At) CK(cudaMalloc(&ix->d_opq_At, (size_t)d * d * sizeof(float)));
  CK(cudaMemcpyAsync(ix->d_opq_At, At.data(), (size_t)d * d * sizeof(float), cudaMemcpyHostToDevice, ix->wstream));
  if (b) {
    if (!ix->d_opq_b) CK(cudaMalloc(&ix->d_opq_b, (size_t)d * sizeof(float)));
    CK(cudaMemcpyAsync(ix->d_opq_b, b, (size_t)d * sizeof(float), cudaMemcpyHostToDevice, ix->wstream));
  } else if (ix->d_opq_b) {
    cudaFree(ix->d_opq_b);
    ix->d_opq_b = nullptr;
  }
  CK(cudaStreamSynchronize(ix->wstream));
  return GB200_OK;
}

// ---- live-docs bitmap: capacity ------------------------------------------------------------
// make d_live hold at least `words` words (new words = all ones).  May replace the array: the caller holds writer_mu and
// NO hold on data_mu.
static int live_reserve(gb200_index *ix, long long words) {
  if (words <= ix->live_words) return GB200_OK;
  ExclusiveScope x(ix);
  long long cap = words + words / 2 + 1024;
  uint32_t *nl = nullptr;
  CK(cudaMalloc(&nl, (size_t)cap * sizeof(uint32_t)));
  CK(cudaMemsetAsync(nl, 0xff, (size_t)cap * sizeof(uint32_t), ix->wstream));
  if (ix->live_words > 0)
    CK(cudaMemcpyAsync(nl, ix->d_live, (size_t)ix->live_words * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ix->wstream));
  CK(cudaStreamSynchronize(ix->wstream));
  if (ix->d_live) cudaFree(ix->d_live);
  ix->d_live = nl;
  ix->live_words = cap;
  return GB200_OK;
}
static int rebuild_installed_filter(gb200_index *ix);  // below
// A writer is about to publish docs up to new_doc_bits: the bitmaps a search may index with those ids must cover them
// first.  Caller holds writer_mu and no hold on data_mu.
static int bitmaps_follow_growth(gb200_index *ix, long long new_doc_bits) {
  if (ix->d_live) CKI(live_reserve(ix, (new_doc_bits + 31) / 32));
  if (ix->inst.active && new_doc_bits > ix->inst.bits) CKI(rebuild_installed_filter(ix));
  return GB200_OK;
}

// ---- posting pools ---------------------------------------------------------------------
// grow the pools to hold need_total postings.  Replaces the arrays: the caller holds data_mu exclusively.
static int pool_reserve_exclusive(gb200_index *ix, long long need_total) {
  if (need_total <= ix->pool_cap) return GB200_OK;
  long long ncap = std::max(need_total, ix->pool_cap * 2);
  ncap = roundup32(ncap);
  const int M = ix->p.nsubvector;
  uint8_t *nc = nullptr;
  int *ni = nullptr;
  float *nn = nullptr;
  if (cudaMalloc(&nc, (size_t)ncap * M) != cudaSuccess || cudaMalloc(&ni, (size_t)ncap * sizeof(int)) != cudaSuccess ||
      cudaMalloc(&nn, (size_t)ncap * sizeof(float)) != cudaSuccess) {
    if (nc) cudaFree(nc);
    if (ni) cudaFree(ni);
    if (nn) cudaFree(nn);
    cudaGetLastError();
    set_err("posting pool: cannot allocate %lld postings", ncap);
    return GB200_ENOMEM;
  }
  if (ix->pool_used > 0) {
    CK(cudaMemcpyAsync(nc, ix->d_codes, (size_t)ix->pool_used * M, cudaMemcpyDeviceToDevice, ix->wstream));
    CK(cudaMemcpyAsync(ni, ix->d_ids, (size_t)ix->pool_used * sizeof(int), cudaMemcpyDeviceToDevice, ix->wstream));
    CK(cudaMemcpyAsync(nn, ix->d_norms, (size_t)ix->pool_used * sizeof(float), cudaMemcpyDeviceToDevice, ix->wstream));
  }
  CK(cudaStreamSynchronize(ix->wstream));
  if (ix->d_codes) cudaFree(ix->d_codes);
  if (ix->d_ids) cudaFree(ix->d_ids);
  if (ix->d_norms) cudaFree(ix->d_norms);
  ix->d_codes = nc;
  ix->d_ids = ni;
  ix->d_norms = nn;
  ix->pool_cap = ncap;
  return GB200_OK;
}

// place n postings (already assigned list/pos) on the device; pos may address existing slots (update in place).
// Regions are resolved through the writer's own offsets table (lists being moved are not published yet).
static int write_postings(gb200_index *ix, long long n, const int *list_no, const int *pos, const int *vid32,
                          const uint8_t *codes) {
  const int M = ix->p.nsubvector;
  size_t b_int = (size_t)n * sizeof(int);
  size_t total = 3 * b_int + (size_t)n * M;
  CKI(ix->w_stage.ensure(total));
  char *base = ix->w_stage.as<char>();
  int *d_list = (int *)base, *d_pos = (int *)(base + b_int), *d_vid = (int *)(base + 2 * b_int);
  uint8_t *d_aos = (uint8_t *)(base + 3 * b_int);
  CK(cudaMemcpyAsync(d_list, list_no, b_int, cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaMemcpyAsync(d_pos, pos, b_int, cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaMemcpyAsync(d_vid, vid32, b_int, cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaMemcpyAsync(d_aos, codes, (size_t)n * M, cudaMemcpyHostToDevice, ix->wstream));
  AppendParams A;
  A.list_no = d_list;
  A.pos = d_pos;
  A.vid = d_vid;
  A.codes_aos = d_aos;
  A.centroids = ix->d_cent;
  A.pq = ix->d_pq;
  A.list_off = ix->d_woff;
  A.codes = ix->d_codes;
  A.ids = ix->d_ids;
  A.norms = ix->d_norms;
  A.n = (int)n;
  A.d = ix->p.d;
  A.M = M;
  A.dsub = ix->dsub;
  A.chunk = ix->chunk;
  A.layout = ix->layout;
  CK(launch_append(A, ix->wstream));
  ix->launches++;
  return GB200_OK;
}

// publish the extents of `lists` (host tables already updated): off first, len with release semantics; the writer's own
// offsets table follows
static int publish_lists(gb200_index *ix, const std::vector<int> &lists) {
  const int n = (int)lists.size();
  if (n == 0) return GB200_OK;
  std::vector<long long> offs(n);
  std::vector<int> lens(n);
  for (int i = 0; i < n; i++) offs[i] = ix->h_off[lists[i]], lens[i] = ix->h_len[lists[i]];
  const size_t bytes = (size_t)n * (sizeof(long long) + 2 * sizeof(int));
  CKI(ix->w_pub.ensure(bytes));
  long long *d_o = ix->w_pub.as<long long>();
  int *d_l = reinterpret_cast<int *>(d_o + n), *d_n = d_l + n;
  CK(cudaMemcpyAsync(d_o, offs.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaMemcpyAsync(d_l, lists.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaMemcpyAsync(d_n, lens.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ix->wstream));
  CK(launch_publish_lists(d_l, d_o, d_n, n, ix->d_off, ix->d_len, ix->wstream));
  ix->launches++;
  CK(cudaStreamSynchronize(ix->wstream));  // the host vectors go out of scope
  return GB200_OK;
}

// caller holds writer_mu and NO hold on data_mu
static int append_locked(gb200_index *ix, int64_t n, const int32_t *list_no, const int64_t *vids,
                         const uint8_t *codes) {
  const int nlist = ix->p.nlist, M = ix->p.nsubvector;
  std::vector<int> inc(nlist, 0);
  long long mv = ix->max_vid.load();
  for (int64_t i = 0; i < n; i++) {
    if (list_no[i] < 0 || list_no[i] >= nlist || vids[i] < 0 || vids[i] > 0x7ffffffeLL) {
      set_err("append: posting %lld has list %d / vid %lld out of range", (long long)i, list_no[i], (long long)vids[i]);
      return GB200_EINVAL;
    }
    inc[list_no[i]]++;
    if (vids[i] > mv) mv = vids[i];
  }
  // growth plan into temporaries: lists that overflow get new contiguous regions at the pool tail.  Nothing of the
  // index is touched until the pool is known to be large enough.
  struct Grow { int list; long long old_off, new_off; int new_cap; };
  std::vector<Grow> grows;
  std::vector<int> touched;
  long long tail = ix->pool_used;
  for (int l = 0; l < nlist; l++) {
    if (!inc[l]) continue;
    touched.push_back(l);
    long long need = (long long)ix->h_len[l] + inc[l];
    if (need > (1LL << GB_SEQ_POS_BITS)) {
      set_err("list %d would hold %lld postings (> 2^21, bucket_max_size)", l, need);
      return GB200_EUNSUPPORTED;
    }
    if (need > ix->h_cap[l]) {
      long long ncap = roundup32(std::max(need, (long long)ix->h_cap[l] + ix->h_cap[l] / 2));
      grows.push_back({l, ix->h_off[l], tail, (int)ncap});
      tail += ncap;
    }
  }
  if (tail > ix->pool_cap) {  // the arrays are replaced: drain the searches for the swap
    ExclusiveScope x(ix);
    CKI(pool_reserve_exclusive(ix, tail));
  }
  CKI(bitmaps_follow_growth(ix, std::max(mv + 1, ix->doc_bits())));
  std::shared_lock<std::shared_mutex> shared(ix->data_mu);
  // commit the plan
  if (tail > ix->pool_used) {
    CK(launch_fill_i32(ix->d_ids + ix->pool_used, tail - ix->pool_used, -1, ix->wstream));
    ix->launches++;
    for (const Grow &g : grows) {
      int len = ix->h_len[g.list];
      ix->pool_live_cap += g.new_cap - ix->h_cap[g.list];
      ix->h_off[g.list] = g.new_off;
      ix->h_cap[g.list] = g.new_cap;
      if (len == 0) continue;
      CK(cudaMemcpyAsync(ix->d_codes + (size_t)g.new_off * M, ix->d_codes + (size_t)g.old_off * M,
                         (size_t)roundup32(len) * M, cudaMemcpyDeviceToDevice, ix->wstream));
      CK(cudaMemcpyAsync(ix->d_ids + g.new_off, ix->d_ids + g.old_off, (size_t)len * sizeof(int),
                         cudaMemcpyDeviceToDevice, ix->wstream));
      CK(cudaMemcpyAsync(ix->d_norms + g.new_off, ix->d_norms + g.old_off, (size_t)len * sizeof(float),
                         cudaMemcpyDeviceToDevice, ix->wstream));
    }
    ix->pool_used = tail;
    // the writer's offsets table sees the new regions now; the published one after the data is in place
    if (grows.size() > 64) {
      CK(cudaMemcpyAsync(ix->d_woff, ix->h_off.data(), sizeof(long long) * nlist, cudaMemcpyHostToDevice, ix->wstream));
    } else {
      for (const Grow &g : grows)
        CK(cudaMemcpyAsync(ix->d_woff + g.list, &ix->h_off[g.list], sizeof(long long), cudaMemcpyHostToDevice, ix->wstream));
    }
  }
  // positions, in arrival order (== list order == tie-break order)
  std::vector<int> pos(n), vid32(n);
  std::vector<int> run(ix->h_len);
  for (int64_t i = 0; i < n; i++) {
    pos[i] = run[list_no[i]]++;
    vid32[i] = (int)vids[i];
    if ((size_t)vids[i] >= ix->vid_loc.size()) ix->vid_loc.resize(std::max<size_t>(vids[i] + 1, ix->vid_loc.size() * 2), -1);
    ix->vid_loc[vids[i]] = ((long long)list_no[i] << 32) | (unsigned)pos[i];
  }
  // stage in slabs so the scratch stays bounded for bulk loads
  const int64_t SLAB = 1 << 22;
  for (int64_t s = 0; s < n; s += SLAB) {
    int64_t m = std::min(SLAB, n - s);
    CKI(write_postings(ix, m, list_no + s, pos.data() + s, vid32.data() + s, codes + (size_t)s * M));
    CK(cudaStreamSynchronize(ix->wstream));  // staging buffer is reused
  }
  ix->h_len = run;
  ix->max_vid = mv;
  {
    int mx = ix->max_list_len.load();
    for (int l : touched) mx = std::max(mx, ix->h_len[l]);
    ix->max_list_len = mx;
  }
  // publish: the scan sees a new length only after the data (and a moved list's new region) are in place
  return publish_lists(ix, touched);
}

int gb200_ivfpq_append(gb200_index *ix, int64_t n, const int32_t *list_no, const int64_t *vids,
                       const uint8_t *codes) {
  if (!ix || ix->kind != 0 || n < 0 || (n > 0 && (!list_no || !vids || !codes))) return GB200_EINVAL;
  if (!ix->trained) {
    set_err("append before set_quantizers");
    return GB200_ENOTTRAINED;
  }
  if (n == 0) return GB200_OK;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  return append_locked(ix, n, list_no, vids, codes);
}

int gb200_ivfpq_update(gb200_index *ix, int64_t vid, int32_t new_list, const uint8_t *code) {
  if (!ix || ix->kind != 0 || !code || new_list < 0 || new_list >= ix->p.nlist) return GB200_EINVAL;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  if (vid < 0 || (size_t)vid >= ix->vid_loc.size() || ix->vid_loc[vid] < 0) return GB200_OK;  // reference: do nothing
  long long loc = ix->vid_loc[vid];
  int old_list = (int)(loc >> 32), old_pos = (int)(loc & 0xffffffff);
  if (old_list == new_list) {
    std::shared_lock<std::shared_mutex> shared(ix->data_mu);
    int l = new_list, p = old_pos, v = (int)vid;
    CKI(write_postings(ix, 1, &l, &p, &v, code));
    CK(cudaStreamSynchronize(ix->wstream));
    return GB200_OK;
  }
  {  // mark the old posting dead: id |= sign bit  (kDelIdxMask analogue)
    std::shared_lock<std::shared_mutex> shared(ix->data_mu);
    int dead = (int)((unsigned)vid | 0x80000000u);
    CK(cudaMemcpyAsync(ix->d_ids + ix->h_off[old_list] + old_pos, &dead, sizeof(int), cudaMemcpyHostToDevice, ix->wstream));
    CK(cudaStreamSynchronize(ix->wstream));
  }
  int64_t v64 = vid;
  return append_locked(ix, 1, &new_list, &v64, code);
}

int gb200_ivfpq_list_sizes(gb200_index *ix, int64_t *sizes) {
  if (!ix || ix->kind != 0 || !sizes) return GB200_EINVAL;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  for (int l = 0; l < ix->p.nlist; l++) sizes[l] = ix->h_len[l];
  return GB200_OK;
}

int gb200_ivfpq_get_list(gb200_index *ix, int32_t list_no, int64_t *ids, uint8_t *codes) {
  if (!ix || ix->kind != 0 || list_no < 0 || list_no >= ix->p.nlist) return GB200_EINVAL;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  std::shared_lock<std::shared_mutex> shared(ix->data_mu);
  int len = ix->h_len[list_no];
  if (len == 0) return GB200_OK;
  const int M = ix->p.nsubvector;
  CKI(ix->w_stage.ensure((size_t)len * (M + sizeof(int))));
  int *d_i = ix->w_stage.as<int>();
  uint8_t *d_c = reinterpret_cast<uint8_t *>(d_i + len);
  CK(launch_gather_list(ix->d_codes, ix->d_ids, ix->h_off[list_no], len, M, ix->chunk, ix->layout, d_c, d_i, ix->wstream));
  ix->launches++;
  std::vector<int> hi(len);
  CK(cudaMemcpyAsync(hi.data(), d_i, (size_t)len * sizeof(int), cudaMemcpyDeviceToHost, ix->wstream));
  CK(cudaMemcpyAsync(codes, d_c, (size_t)len * M, cudaMemcpyDeviceToHost, ix->wstream));
  CK(cudaStreamSynchronize(ix->wstream));
  for (int i = 0; i < len; i++) {
    unsigned u = (unsigned)hi[i];
    ids[i] = (u & 0x80000000u) ? (int64_t)((uint64_t)(u & 0x7fffffffu) | 0x8000000000000000ull) : (int64_t)u;
  }
  return GB200_OK;
}

// Device-side compaction (RealTimeMemData::CompactBucket, realtime_mem_data.cc:354-424): drop moved (kDelIdxMask) and
// bitmap-deleted postings, keep the survivors' order.
//   list_no >= 0: that list is rewritten into a fresh region at the pool tail and swapped in by publication — searches
//                 keep running (they see the old or the new extent, both consistent);
//   list_no == -1: every list is compacted into a NEW, tightly packed pool (this also returns the regions abandoned
//                 by list growth); the pools are replaced, so searches are drained for the swap.
int gb200_ivfpq_compact(gb200_index *ix, int32_t list_no, int64_t *dropped) {
  if (!ix || ix->kind != 0 || list_no < -1 || list_no >= ix->p.nlist) return GB200_EINVAL;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  const int nlist = ix->p.nlist, M = ix->p.nsubvector;
  const int n_lists = list_no >= 0 ? 1 : nlist;
  if (dropped) *dropped = 0;
  if (ix->pool_used == 0) return GB200_OK;
  std::vector<int> lists(n_lists), new_len(n_lists), new_cap(n_lists);
  std::vector<long long> new_off(n_lists);
  for (int i = 0; i < n_lists; i++) lists[i] = list_no >= 0 ? list_no : i;
  const size_t st_bytes = (size_t)n_lists * (sizeof(long long) + 3 * sizeof(int));
  CompactParams C;
  memset(&C, 0, sizeof(C));
  C.M = M;
  C.chunk = ix->chunk;
  C.layout = ix->layout;
  auto fill_sources = [&](CompactParams &c) {
    long long *d_noff = ix->w_pub.as<long long>();
    int *d_lists = reinterpret_cast<int *>(d_noff + n_lists);
    c.lists = d_lists;
    c.new_off = d_noff;
    c.new_len = d_lists + n_lists;
    c.new_cap = d_lists + 2 * n_lists;
    c.list_off = ix->d_off;
    c.list_len = ix->d_len;
    c.codes = ix->d_codes;
    c.ids = ix->d_ids;
    c.norms = ix->d_norms;
    c.live = ix->deleted_count.load() > 0 ? ix->d_live : nullptr;
    c.live_bits = ix->live_words * 32;
  };
  {  // pass 0: survivors per list
    std::shared_lock<std::shared_mutex> shared(ix->data_mu);
    CKI(ix->w_pub.ensure(st_bytes));
    fill_sources(C);
    CK(cudaMemcpyAsync(const_cast<int *>(C.lists), lists.data(), (size_t)n_lists * sizeof(int), cudaMemcpyHostToDevice, ix->wstream));
    CK(launch_compact_lists(C, n_lists, ix->wstream));
    ix->launches++;
    CK(cudaMemcpyAsync(new_len.data(), C.new_len, (size_t)n_lists * sizeof(int), cudaMemcpyDeviceToHost, ix->wstream));
    CK(cudaStreamSynchronize(ix->wstream));
  }
  long long drop = 0;
  for (int i = 0; i < n_lists; i++) drop += ix->h_len[lists[i]] - new_len[i];
  if (dropped) *dropped = drop;
  if (drop == 0 && list_no >= 0) return GB200_OK;
  // destination regions
  long long tail = list_no >= 0 ? ix->pool_used : 0;
  for (int i = 0; i < n_lists; i++) {
    // single list: keep the old capacity, so that a search still holding the old (longer) length stays inside the
    // region; full rebuild: tight, with the usual head-room
    new_cap[i] = list_no >= 0 ? std::max(ix->h_cap[lists[i]], (int)roundup32(new_len[i]))
                              : (int)roundup32(new_len[i] + new_len[i] / 8);
    new_off[i] = tail;
    tail += new_cap[i];
  }
  std::unique_ptr<ExclusiveScope> excl;
  std::shared_lock<std::shared_mutex> shared(ix->data_mu, std::defer_lock);
  uint8_t *nc = nullptr;
  int *ni = nullptr;
  float *nn = nullptr;
  long long ncap_pool = ix->pool_cap;
  if (list_no >= 0) {
    if (tail > ix->pool_cap) {
      ExclusiveScope x(ix);
      CKI(pool_reserve_exclusive(ix, tail));
    }
    shared.lock();
    nc = ix->d_codes, ni = ix->d_ids, nn = ix->d_norms;
  } else {
    excl.reset(new ExclusiveScope(ix));
    ncap_pool = roundup32(std::max<long long>(tail + tail / 4, 1024));
    if (cudaMalloc(&nc, (size_t)ncap_pool * M) != cudaSuccess || cudaMalloc(&ni, (size_t)ncap_pool * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&nn, (size_t)ncap_pool * sizeof(float)) != cudaSuccess) {
      if (nc) cudaFree(nc);
      if (ni) cudaFree(ni);
      if (nn) cudaFree(nn);
      cudaGetLastError();
      set_err("compaction: cannot allocate the new pool (%lld postings)", ncap_pool);
      return GB200_ENOMEM;
    }
  }
  // pass 1: move the survivors
  fill_sources(C);
  CK(cudaMemcpyAsync(const_cast<long long *>(C.new_off), new_off.data(), (size_t)n_lists * sizeof(long long), cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaMemcpyAsync(const_cast<int *>(C.new_cap), new_cap.data(), (size_t)n_lists * sizeof(int), cudaMemcpyHostToDevice, ix->wstream));
  C.dst_codes = nc;
  C.dst_ids = ni;
  C.dst_norms = nn;
  CK(launch_compact_lists(C, n_lists, ix->wstream));
  ix->launches++;
  CK(cudaStreamSynchronize(ix->wstream));
  // host tables, vid -> (list, pos) of the survivors
  std::vector<int> hid;
  for (int i = 0; i < n_lists; i++) {
    const int l = lists[i];
    if (list_no >= 0) ix->pool_live_cap += new_cap[i] - ix->h_cap[l];
    ix->h_off[l] = new_off[i];
    ix->h_len[l] = new_len[i];
    ix->h_cap[l] = new_cap[i];
    if (new_len[i] == 0) continue;
    hid.resize(new_len[i]);
    CK(cudaMemcpy(hid.data(), ni + new_off[i], (size_t)new_len[i] * sizeof(int), cudaMemcpyDeviceToHost));
    for (int p = 0; p < new_len[i]; p++)
      if (hid[p] >= 0 && (size_t)hid[p] < ix->vid_loc.size()) ix->vid_loc[hid[p]] = ((long long)l << 32) | (unsigned)p;
  }
  if (list_no >= 0) {
    ix->pool_used = tail;
    CK(cudaMemcpyAsync(ix->d_woff + list_no, &ix->h_off[list_no], sizeof(long long), cudaMemcpyHostToDevice, ix->wstream));
    return publish_lists(ix, lists);
  }
  // full rebuild: swap the pools (searches are drained), publish every extent
  cudaFree(ix->d_codes);
  cudaFree(ix->d_ids);
  cudaFree(ix->d_norms);
  ix->d_codes = nc;
  ix->d_ids = ni;
  ix->d_norms = nn;
  ix->pool_cap = ncap_pool;
  ix->pool_used = tail;
  ix->pool_live_cap = tail;
  CK(launch_fill_i32(ix->d_ids + tail, ncap_pool - tail, -1, ix->wstream));
  CK(cudaMemcpyAsync(ix->d_off, ix->h_off.data(), sizeof(long long) * nlist, cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaMemcpyAsync(ix->d_woff, ix->h_off.data(), sizeof(long long) * nlist, cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaMemcpyAsync(ix->d_len, ix->h_len.data(), sizeof(int) * nlist, cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaStreamSynchronize(ix->wstream));
  return GB200_OK;
}

int gb200_ivfpq_replace_list(gb200_index *ix, int32_t list_no, int64_t n, const int64_t *ids, const uint8_t *codes) {
  if (!ix || ix->kind != 0 || list_no < 0 || list_no >= ix->p.nlist || n < 0 || (n > 0 && (!ids || !codes)))
    return GB200_EINVAL;
  if (!ix->trained) return GB200_ENOTTRAINED;
  if (n > (1LL << GB_SEQ_POS_BITS)) return GB200_EUNSUPPORTED;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  const int M = ix->p.nsubvector, l = list_no;
  long long mv = ix->max_vid.load();
  for (int64_t i = 0; i < n; i++) {
    const int64_t v = ids[i] & 0x7fffffffffffffffLL;
    if (v > 0x7ffffffeLL) {
      set_err("replace_list: posting %lld has vid %lld out of range", (long long)i, (long long)v);
      return GB200_EINVAL;
    }
    if (ids[i] >= 0 && v > mv) mv = v;
  }
  // vids that live in the old content lose their location
  const int old_len = ix->h_len[l];
  if (old_len > 0) {
    std::vector<int> old(old_len);
    std::shared_lock<std::shared_mutex> shared(ix->data_mu);
    CK(cudaMemcpyAsync(old.data(), ix->d_ids + ix->h_off[l], (size_t)old_len * sizeof(int), cudaMemcpyDeviceToHost, ix->wstream));
    CK(cudaStreamSynchronize(ix->wstream));
    for (int p = 0; p < old_len; p++)
      if (old[p] >= 0 && (size_t)old[p] < ix->vid_loc.size() && ix->vid_loc[old[p]] == (((long long)l << 32) | (unsigned)p))
        ix->vid_loc[old[p]] = -1;
  }
  const long long new_cap = roundup32(std::max<long long>(n + n / 8 + 1, 32));
  const long long off = ix->pool_used, tail = off + new_cap;
  if (tail > ix->pool_cap) {
    ExclusiveScope x(ix);
    CKI(pool_reserve_exclusive(ix, tail));
  }
  CKI(bitmaps_follow_growth(ix, std::max(mv + 1, ix->doc_bits())));
  std::shared_lock<std::shared_mutex> shared(ix->data_mu);
  CK(launch_fill_i32(ix->d_ids + off, new_cap, -1, ix->wstream));
  ix->launches++;
  // the writer's offsets table sees the new region now, the published one after the data is in place
  CK(cudaMemcpyAsync(ix->d_woff + l, &off, sizeof(long long), cudaMemcpyHostToDevice, ix->wstream));
  CK(cudaStreamSynchronize(ix->wstream));  // &off is a stack address
  if (n > 0) {
    std::vector<int> lno((size_t)n, l), pos((size_t)n), vid32((size_t)n);
    for (int64_t i = 0; i < n; i++) {
      const int v = (int)(ids[i] & 0x7fffffffLL);
      pos[i] = (int)i;
      if (ids[i] < 0) {  // kDelIdxMask: the slot stays, dead
        vid32[i] = (int)((unsigned)v | 0x80000000u);
        continue;
      }
      vid32[i] = v;
      if ((size_t)v >= ix->vid_loc.size()) ix->vid_loc.resize(std::max<size_t>((size_t)v + 1, ix->vid_loc.size() * 2), -1);
      const long long prev = ix->vid_loc[v];
      if (prev >= 0 && (int)(prev >> 32) != l) {  // still alive in another list on the device: a vid lives in one place
        const int dead = (int)((unsigned)v | 0x80000000u);
        CK(cudaMemcpyAsync(ix->d_ids + ix->h_off[(int)(prev >> 32)] + (int)(prev & 0xffffffff), &dead, sizeof(int),
                           cudaMemcpyHostToDevice, ix->wstream));
        CK(cudaStreamSynchronize(ix->wstream));
      }
      ix->vid_loc[v] = ((long long)l << 32) | (unsigned)i;
    }
    const int64_t SLAB = 1 << 22;
    for (int64_t s0 = 0; s0 < n; s0 += SLAB) {
      const int64_t m = std::min(SLAB, n - s0);
      CKI(write_postings(ix, m, lno.data() + s0, pos.data() + s0, vid32.data() + s0, codes + (size_t)s0 * M));
      CK(cudaStreamSynchronize(ix->wstream));
    }
  }
  ix->pool_live_cap += new_cap - ix->h_cap[l];
  ix->h_off[l] = off;
  ix->h_cap[l] = (int)new_cap;
  ix->h_len[l] = (int)n;
  ix->pool_used = tail;
  ix->max_vid = mv;
  ix->max_list_len = std::max(ix->max_list_len.load(), (int)n);
  return publish_lists(ix, std::vector<int>(1, l));
}

// ---- raw vectors -----------------------------------------------------------------------
static int upload_raw_impl(gb200_index *ix, int64_t first_vid, int64_t n, const float *x, cudaMemcpyKind kind) {
  if (!ix || first_vid < 0 || n < 0 || (n > 0 && !x)) return GB200_EINVAL;
  if (n == 0) return GB200_OK;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  const int rd = ix->p.raw_d;
  long long need = first_vid + n;
  if (need > 0x7fffffffLL) return GB200_EUNSUPPORTED;
  const long long have = ix->raw_n.load();
  if (need > ix->raw_cap) {  // the store is replaced: drain the searches for the swap
    ExclusiveScope xs(ix);
    long long ncap = std::max(need, ix->raw_cap + ix->raw_cap / 2);
    float *nr = nullptr;
    CK(cudaMalloc(&nr, (size_t)ncap * rd * sizeof(float)));
    if (have > 0)
      CK(cudaMemcpyAsync(nr, ix->d_raw, (size_t)have * rd * sizeof(float), cudaMemcpyDeviceToDevice, ix->wstream));
    CK(cudaStreamSynchronize(ix->wstream));
    if (ix->d_raw) cudaFree(ix->d_raw);
    ix->d_raw = nr;
    ix->raw_cap = ncap;
  }
  if (need > have) CKI(bitmaps_follow_growth(ix, std::max(need, ix->doc_bits())));
  std::shared_lock<std::shared_mutex> shared(ix->data_mu);
  if (first_vid > have)  // rows nobody uploaded (a gap) read as zero vectors rather than as whatever the allocation held
    CK(cudaMemsetAsync(ix->d_raw + (size_t)have * rd, 0, (size_t)(first_vid - have) * rd * sizeof(float), ix->wstream));
  CK(cudaMemcpyAsync(ix->d_raw + (size_t)first_vid * rd, x, (size_t)n * rd * sizeof(float), kind, ix->wstream));
  CK(cudaStreamSynchronize(ix->wstream));
  {
    std::lock_guard<std::mutex> a(ix->aux_mu);
    if (first_vid < ix->aux_n.load()) ix->aux_n = first_vid;  // companions of the rewritten rows are stale
  }
  if (need > have) ix->raw_n = need;  // published after the rows are in place
  return GB200_OK;
}

int gb200_upload_raw(gb200_index *ix, int64_t first_vid, int64_t n, const float *x) {
  return upload_raw_impl(ix, first_vid, n, x, cudaMemcpyHostToDevice);
}
int gb200_upload_raw_dev(gb200_index *ix, int64_t first_vid, int64_t n, const float *x_dev) {
  return upload_raw_impl(ix, first_vid, n, x_dev, cudaMemcpyDeviceToDevice);
}

int64_t gb200_raw_count(gb200_index *ix) { return ix ? ix->raw_n.load() : 0; }

// ---- live-docs bitmap: contents ----------------------------------------------------------------------
// upload the words of h_deleted listed in `touched` (as ~deleted) — only what changed travels
static int push_live_words(gb200_index *ix, const std::vector<long long> &touched) {
  if (touched.empty()) return GB200_OK;
  const int n = (int)touched.size();
  std::vector<uint32_t> vals(n);
  for (int i = 0; i < n; i++) vals[i] = ~ix->h_deleted[(size_t)touched[i]];
  CKI(live_reserve(ix, std::max<long long>((long long)ix->h_deleted.size(), (ix->doc_bits() + 31) / 32)));
  {
    std::shared_lock<std::shared_mutex> shared(ix->data_mu);
    CKI(ix->w_stage.ensure((size_t)n * (sizeof(long long) + sizeof(uint32_t))));
    long long *d_idx = ix->w_stage.as<long long>();
    uint32_t *d_val = reinterpret_cast<uint32_t *>(d_idx + n);
    CK(cudaMemcpyAsync(d_idx, touched.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, ix->wstream));
    CK(cudaMemcpyAsync(d_val, vals.data(), (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->wstream));
    CK(launch_scatter_words(d_idx, d_val, n, ix->d_live, ix->wstream));
    ix->launches++;
    CK(cudaStreamSynchronize(ix->wstream));
  }
  return GB200_OK;
}

int gb200_set_deleted(gb200_index *ix, const int64_t *docids, int64_t n, int deleted) {
  if (!ix || n < 0 || (n > 0 && !docids)) return GB200_EINVAL;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  for (int64_t i = 0; i < n; i++)
    if (docids[i] < 0 || docids[i] > 0x7ffffffeLL) return GB200_EINVAL;
  std::vector<long long> touched;
  long long cnt = ix->deleted_count.load();
  for (int64_t i = 0; i < n; i++) {
    int64_t doc = docids[i];
    size_t wd = (size_t)(doc >> 5);
    if (wd >= ix->h_deleted.size()) ix->h_deleted.resize(std::max(wd + 1, ix->h_deleted.size() * 2), 0u);
    const uint32_t bit = 1u << (doc & 31), old = ix->h_deleted[wd];
    const uint32_t nw = deleted ? (old | bit) : (old & ~bit);
    if (nw != old) {
      ix->h_deleted[wd] = nw;
      cnt += deleted ? 1 : -1;
      touched.push_back((long long)wd);
    }
  }
  std::sort(touched.begin(), touched.end());
  touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
  CKI(push_live_words(ix, touched));
  ix->deleted_count = cnt;
  if (!touched.empty() && ix->inst.active) CKI(rebuild_installed_filter(ix));  // the installed filter folds the deleted bits in
  return GB200_OK;
}

// Bring the device bitmap in line with the reference's BitmapManager contents (bit = 1: deleted).  Only words that differ
// from the shadow copy are uploaded, so calling this before every Search (the reference tests the bitmap live, and
// some engine paths set bits without calling RetrievalModel::Delete — search/gamma_engine.cc:866) costs one memcmp.
int gb200_upload_deleted_bitmap(gb200_index *ix, const uint8_t *bitmap, int64_t nbits) {
  if (!ix || nbits < 0 || (nbits > 0 && !bitmap)) return GB200_EINVAL;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  const size_t words = (size_t)((nbits + 31) / 32), bytes = (size_t)((nbits + 7) / 8);
  if (words > ix->h_deleted.size()) ix->h_deleted.resize(words, 0u);
  // byte[id>>3] & (1 << (id&7))  ==  little-endian u32 word[id>>5] bit (id&31)
  std::vector<long long> touched;
  long long cnt = ix->deleted_count.load();
  const size_t full = bytes / 4;
  const uint8_t *shadow = reinterpret_cast<const uint8_t *>(ix->h_deleted.data());
  for (size_t w0 = 0; w0 < full; w0 += 1024) {  // compare 4 KB at a time, look closer only where something changed
    const size_t w1 = std::min(full, w0 + 1024);
    if (!memcmp(shadow + w0 * 4, bitmap + w0 * 4, (w1 - w0) * 4)) continue;
    for (size_t wd = w0; wd < w1; wd++) {
      uint32_t nw;
      memcpy(&nw, bitmap + wd * 4, 4);
      if (nw != ix->h_deleted[wd]) {
        cnt += __builtin_popcount(nw) - __builtin_popcount(ix->h_deleted[wd]);
        ix->h_deleted[wd] = nw;
        touched.push_back((long long)wd);
      }
    }
  }
  if (full < words) {  // last, partial word
    uint32_t nw = 0;
    memcpy(&nw, bitmap + full * 4, bytes - full * 4);
    if (nbits & 31) nw &= (1u << (nbits & 31)) - 1u;
    if (nw != ix->h_deleted[full]) {
      cnt += __builtin_popcount(nw) - __builtin_popcount(ix->h_deleted[full]);
      ix->h_deleted[full] = nw;
      touched.push_back((long long)full);
    }
  }
  for (size_t wd = words; wd < ix->h_deleted.size(); wd++)  // docs beyond the given bitmap are not deleted
    if (ix->h_deleted[wd]) {
      cnt -= __builtin_popcount(ix->h_deleted[wd]);
      ix->h_deleted[wd] = 0;
      touched.push_back((long long)wd);
    }
  CKI(push_live_words(ix, touched));
  ix->deleted_count = cnt;
  if (!touched.empty() && ix->inst.active) CKI(rebuild_installed_filter(ix));
  return GB200_OK;
}

// ---- validity bitmap of one search: live AND all range filters -------------------------------------------------
static int build_filter_bitmap(gb200_index *ix, const gb200_range_filter *filters, int n_filters, long long bits,
                               DevBuf &out, DevBuf &fbytes, DevBuf &fdesc, cudaStream_t st) {
  std::vector<DevRangeFilter> desc(n_filters);
  size_t total = 0;
  std::vector<size_t> offs(n_filters);
  for (int f = 0; f < n_filters; f++) {
    const gb200_range_filter &rf = filters[f];
    if (!rf.bitmap || rf.max_doc < rf.min_doc || rf.min_aligned > rf.min_doc || rf.min_aligned < 0) {
      set_err("range filter %d malformed", f);
      return GB200_EINVAL;
    }
    long long max_aligned = ((long long)rf.max_doc / 8 + 1) * 8 - 1;
    size_t nbytes = (size_t)((max_aligned - rf.min_aligned + 1) / 8);
    offs[f] = total;
    total += (nbytes + 15) & ~(size_t)15;
  }
  CKI(fbytes.ensure(total));
  CKI(fdesc.ensure(sizeof(DevRangeFilter) * n_filters));
  for (int f = 0; f < n_filters; f++) {
    const gb200_range_filter &rf = filters[f];
    long long max_aligned = ((long long)rf.max_doc / 8 + 1) * 8 - 1;
    size_t nbytes = (size_t)((max_aligned - rf.min_aligned + 1) / 8);
    CK(cudaMemcpyAsync(fbytes.as<uint8_t>() + offs[f], rf.bitmap, nbytes, cudaMemcpyHostToDevice, st));
    desc[f].bitmap = fbytes.as<uint8_t>() + offs[f];
    desc[f].min_doc = rf.min_doc;
    desc[f].max_doc = rf.max_doc;
    desc[f].min_aligned = rf.min_aligned;
    desc[f].not_in = rf.not_in;
  }
  CK(cudaMemcpyAsync(fdesc.p, desc.data(), sizeof(DevRangeFilter) * n_filters, cudaMemcpyHostToDevice, st));
  CKI(out.ensure((size_t)bits / 8));
  CK(launch_build_valid(ix->deleted_count.load() > 0 ? ix->d_live : nullptr, ix->live_words * 32,
                        fdesc.as<DevRangeFilter>(), n_filters, out.as<uint32_t>(), bits, st));
  CK(cudaStreamSynchronize(st));  // desc / the caller's filter bytes may go away
  return GB200_OK;
}

// returns device pointer (or nullptr = everything valid) and the number of docs it covers
static int prepare_valid(gb200_index *ix, SearchCtx &c, const gb200_range_filter *filters, int n_filters,
                         const uint32_t **out, long long *out_bits) {
  *out = nullptr;
  *out_bits = 0;
  if (n_filters <= 0) {
    if (ix->deleted_count.load() <= 0) return GB200_OK;
    *out = ix->d_live;
    *out_bits = ix->live_words * 32;
    return GB200_OK;
  }
  const long long bits = roundup32(std::max<long long>(ix->doc_bits(), 32));
  CKI(build_filter_bitmap(ix, filters, n_filters, bits, c.valid_filt, c.filt_bytes, c.filt_desc, c.stream));
  c.launches++;
  *out = c.valid_filt.as<uint32_t>();
  *out_bits = bits;
  return GB200_OK;
}

// (re)build the installed filter's bitmap from its host copy: covers every doc known now plus head-room, so that
// postings appended later fall inside it (docs outside a range's [min, max] fail or pass by its own rule).
// Caller holds writer_mu and no hold on data_mu.
static int rebuild_installed_filter(gb200_index *ix) {
  gb200_index::Installed &I = ix->inst;
  std::vector<gb200_range_filter> f(I.desc);
  long long top = ix->doc_bits();
  for (size_t i = 0; i < f.size(); i++) {
    f[i].bitmap = I.bytes[i].data();
    top = std::max<long long>(top, (long long)f[i].max_doc + 1);
  }
  const long long bits = roundup32(std::max<long long>(top + top / 4, 1024));
  ExclusiveScope x(ix);  // the bitmap the *_dev searches read is rewritten (and possibly reallocated)
  CKI(build_filter_bitmap(ix, f.data(), (int)f.size(), bits, I.valid, I.filt_bytes, I.filt_desc, ix->wstream));
  ix->launches++;
  I.bits = bits;
  return GB200_OK;
}

int gb200_set_filters(gb200_index *ix, const gb200_range_filter *filters, int n_filters) {
  if (!ix) return GB200_EINVAL;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  CKI(use_device(ix));
  gb200_index::Installed &I = ix->inst;
  if (n_filters <= 0) {
    ExclusiveScope x(ix);
    I.active = false;
    I.desc.clear();
    I.bytes.clear();
    return GB200_OK;
  }
  if (!filters) return GB200_EINVAL;
  std::vector<gb200_range_filter> desc(filters, filters + n_filters);
  std::vector<std::vector<uint8_t>> bytes(n_filters);
  for (int f = 0; f < n_filters; f++) {
    const gb200_range_filter &rf = filters[f];
    if (!rf.bitmap || rf.max_doc < rf.min_doc || rf.min_aligned > rf.min_doc || rf.min_aligned < 0) {
      set_err("range filter %d malformed", f);
      return GB200_EINVAL;
    }
    long long max_aligned = ((long long)rf.max_doc / 8 + 1) * 8 - 1;
    bytes[f].assign(rf.bitmap, rf.bitmap + (size_t)((max_aligned - rf.min_aligned + 1) / 8));
  }
  I.desc.swap(desc);
  I.bytes.swap(bytes);
  int rc = rebuild_installed_filter(ix);
  I.active = rc == GB200_OK;
  return rc;
}

// ---- search ------------------------------------------------------------------------------------
static int resolve_nprobe(gb200_index *ix, const gb200_search_params *sp, int dflt) {
  int np = sp->nprobe;
  if (np <= 0 || np > ix->p.nlist) np = dflt;  // gamma_index_ivfpq.cc:539-545
  if (np > ix->p.nlist) np = ix->p.nlist;
  return np;
}

static int coarse_dev(gb200_index *ix, SearchCtx &c, int n, const float *d_xq, int nprobe, int *d_keys, float *d_cdis) {
  const int d = ix->p.d, nlist = ix->p.nlist;
  if (nprobe > 2048) {
    set_err("nprobe=%d > 2048 not implemented", nprobe);
    return GB200_EUNSUPPORTED;
  }
  CKI(c.ws_xn.ensure((size_t)n * sizeof(float)));
  // distance producer: tcgen05 3xTF32 GEMM (default) or the CUDA-core fp32 kernel (GB200_COARSE=simt)
  const bool use_tc = !ix->tune.coarse_simt && (d % 4 == 0) && (nlist % 4 == 0);
  if (use_tc) {
    CKI(c.ws_xs.ensure((size_t)n * d * sizeof(float)));
    if (c.zero_req == n && c.ws_ctl.p) {  // the scan's control words and counter ride on this launch (no memset nodes)
      CK(launch_rows_prep(d_xq, n, d, c.ws_xn.as<float>(), c.ws_xs.as<float>(), c.stream, c.ws_ctl.as<int>(), 4 + 2 * n,
                          c.d_scanned));
      c.zeroed_for = n;
    } else {
      CK(launch_rows_prep(d_xq, n, d, c.ws_xn.as<float>(), c.ws_xs.as<float>(), c.stream));
    }
  } else {
    CK(launch_row_norms(d_xq, n, d, c.ws_xn.as<float>(), c.stream));
  }
  // bound the distance matrix scratch to ~1 GiB by chunking the queries
  long long rows = std::max<long long>(1, (1LL << 28) / nlist);
  if (rows > n) rows = n;
  CKI(c.ws_dist.ensure((size_t)rows * nlist * sizeof(float)));
  // the GEMM epilogue also emits the minimum of every 32-centroid chunk; the select then reads ~nprobe chunks of a row
  // instead of all nlist distances (coarse.cu coarse_select_cmin_kernel)
  const bool use_cmin = use_tc && !ix->tune.coarse_full_select && coarse_select_cmin_usable(nlist, nprobe);
  const int cmin_pitch = tc_gemm_cmin_pitch(nlist);
  if (use_cmin) CKI(c.ws_cmin.ensure((size_t)rows * cmin_pitch * sizeof(float)));
  for (long long r0 = 0; r0 < n; r0 += rows) {
    int m = (int)std::min<long long>(rows, n - r0);
    if (use_tc) {
      CK(launch_tc_gemm(d_xq + (size_t)r0 * d, c.ws_xs.as<float>() + (size_t)r0 * d, c.ws_xn.as<float>() + r0,
                        ix->d_cent, ix->d_cent_small, ix->d_cent_norm, m, nlist, d, c.ws_dist.as<float>(), nlist, 1,
                        use_cmin ? c.ws_cmin.as<float>() : nullptr, cmin_pitch, c.stream));
    } else {
      CK(launch_coarse_dist(d_xq + (size_t)r0 * d, c.ws_xn.as<float>() + r0, ix->d_cent, ix->d_cent_norm, m, nlist, d,
                            c.ws_dist.as<float>(), c.stream));
    }
    if (use_cmin)
      CK(launch_coarse_select_cmin(c.ws_dist.as<float>(), c.ws_cmin.as<float>(), cmin_pitch, m, nlist, nprobe,
                                   d_keys + (size_t)r0 * nprobe, d_cdis + (size_t)r0 * nprobe, c.stream));
    else
      CK(launch_coarse_select(c.ws_dist.as<float>(), m, nlist, nprobe, d_keys + (size_t)r0 * nprobe,
                              d_cdis + (size_t)r0 * nprobe, c.stream));
    c.launches += 2;
  }
  c.launches += 1;
  return GB200_OK;
}

// Positional work plan of one v2 / M = 64 scan launch: queries [0, n_full) are one work item each,
// queries [n_full, n) are s_tail items each, rows = candidate rows per query in the [n][rows][R] buffer.
// n >= slots: whole waves of resident CTAs stay unsplit, the last partial wave is split slots / n_tail ways;
// n < slots: every query is split s_uniform ways (the caller's whole-wave heuristic).  Pure arithmetic (CPU-testable
// through gb200_debug_plan).
static void plan_work(int n, int slots, int nprobe, int R, int s_uniform, int tail_override, int *n_full, int *s_tail,
                      int *n_items, int *rows) {
  int nf = 0, st = 1;
  if (n >= slots) {
    nf = (n / slots) * slots;
    const int n_tail = n - nf;
    st = n_tail ? std::max(1, std::min(std::min(8, nprobe), slots / n_tail)) : 1;
  } else {
    st = s_uniform;
  }
  if (tail_override > 0) st = tail_override;
  st = std::max(1, std::min(st, nprobe));
  while (st > 1 && (long long)st * R > 8192) st--;
  *n_full = nf;
  *s_tail = st;
  *n_items = nf + (n - nf) * st;
  *rows = (nf == n) ? 1 : st;
}

int gb200_debug_plan(int n, int slots, int nprobe, int recall_num, int s_uniform, int *out4) {
  if (n <= 0 || slots <= 0 || nprobe <= 0 || recall_num <= 0 || s_uniform <= 0 || !out4) return GB200_EINVAL;
  plan_work(n, slots, nprobe, recall_num, s_uniform, 0, &out4[0], &out4[1], &out4[2], &out4[3]);
  return GB200_OK;
}

// scan + rerank with probes already on the device
// d_xq: the queries the quantizers see (OPQ applied if the model has one); d_xq_raw: as given, for the exact re-rank
static int scan_rerank_dev(gb200_index *ix, SearchCtx &c, int n, const float *d_xq, const float *d_xq_raw, int k,
                           const gb200_search_params *sp, int nprobe, const int *d_keys, const float *d_cdis,
                           const uint32_t *d_valid, long long valid_bits, float *d_out_d, long long *d_out_i) {
  const int M = ix->p.nsubvector;
  const Tuning &T = ix->tune;
  int R = sp->recall_num < k ? k : sp->recall_num;  // gamma_index_ivfpq.cc:762-765
  const bool ip = sp->metric == GB200_METRIC_INNER_PRODUCT;
  if (R > 2048) {
    set_err("recall_num=%d > 2048 not implemented", R);
    return GB200_EUNSUPPORTED;
  }
  if (nprobe > 2048) return GB200_EUNSUPPORTED;
  if (sp->has_rank && (!ix->d_raw)) {
    set_err("has_rank search but no raw vectors uploaded");
    return GB200_EINVAL;
  }
  // ---- kernel, CTA shape, candidate buffer.
  //  mode 1 (M = 32): persistent kernel with dynamic items (variant 3, ivfpq_scan_v3.cu);
  //  mode 2 (M = 64): one CTA per (query, split), 384 threads, 2 CTAs per SM, positional work plan (variant 2);
  //  mode 0: generic kernel (any M % 4 == 0).
  const int variant = ix->mode == 1 ? 3 : ix->mode == 2 ? 2 : 0;
  int threads = T.scan_threads;
  int cap = scan_buffer_cap(R);  // power of two >= R + 512, >= 1024
  int ctas_per_sm = 3;
  if (ix->mode == 1) {
    if (threads == 0) threads = 384;
    // tuning: a larger buffer means fewer in-loop selects per query (any multiple of 32 the 4-keys-per-thread select holds)
    if (T.scan_cap && cap <= 1024 && T.scan_cap >= R + threads && T.scan_cap <= 4 * threads) cap = T.scan_cap & ~31;
    if (cap > 4 * threads) threads = 512;          // R > 512: 2048 / 4096 keys (4 / 8 per thread in the select)
    if (threads == 512 && cap < 2048) cap = 2048;  // room for one block of each of the 16 warps
    ctas_per_sm = scan_v3_ctas_per_sm(threads, cap);
  } else if (ix->mode == 2) {
    threads = 384;
    if (cap < R + 384) cap *= 2;  // room for one block of every warp above the survivors
    ctas_per_sm = 2;
  }
  const int slots = ix->num_sms * ctas_per_sm;

  ScanParams P;
  memset(&P, 0, sizeof(P));
  P.xq = d_xq;
  P.keys = d_keys;
  P.coarse_dis = d_cdis;
  P.centroids = ix->d_cent;
  P.pq_t = ix->d_pq_t;
  P.codes = ix->d_codes;
  P.ids = ix->d_ids;
  P.norms = ix->d_norms;
  P.list_off = ix->d_off;
  P.list_len = ix->d_len;
  P.valid = d_valid;
  P.valid_bits = valid_bits;
  P.scanned = c.d_scanned;
  P.n = n;
  P.d = ix->p.d;
  P.M = M;
  P.dsub = ix->dsub;
  P.nlist = ix->p.nlist;
  P.nprobe = nprobe;
  P.R = R;
  P.cap = cap;
  P.chunk = ix->chunk;
  P.is_ip = ip ? 1 : 0;
  P.m32_threads = threads;
  P.variant = variant;
  P.pf_blocks = T.pf_blocks;
  P.s_tail = 1;

  // ---- how the batch is cut into CTAs
  int S = 1;  // candidate rows per query in the [n][S][R] buffer
  bool plan = false;
  int n_full = 0, s_tail = 1, n_items = 0, grid_v3 = 0;
  if (variant == 3) {
    // every query may be worked on by up to S CTAs at once (idle CTAs join running queries, ivfpq_scan_v3.cu)
    S = T.max_rows > 0 ? T.max_rows : (n >= slots ? 4 : std::min(8, (slots + n - 1) / n));
    S = std::max(1, std::min(S, 16));
    while (S > 1 && (long long)S * R > 8192) S--;
    grid_v3 = (int)std::min<long long>(slots, (long long)n * S);
    // per-query item table (item -> list, blocks): nprobe x items of the longest list bounds it; coarser items when a
    // few huge lists would make it large (items beyond the table are still found, by search)
    int ch = T.ch_blocks;
    const long long max_blocks = std::max(1, (ix->max_list_len.load() + 31) / 32);
    while ((long long)nprobe * ((max_blocks + ch - 1) / ch) > 512 && ch < 32768) ch *= 2;
    P.v3_max_items = (int)std::min<long long>(512, std::max<long long>(2, (long long)nprobe * ((max_blocks + ch - 1) / ch)));
    P.ch_blocks = ch;
    P.help_min = T.help_min;
    P.v3_tma = T.v3_tma;
    P.v3_flags = T.v3_flags;
    P.help_window = std::min(n, 4 * slots);
    P.max_np_s = nprobe;
  } else {
    // one CTA per (query, split), probes dealt round-robin.  Uniform S: the value in 1..8 that best fills whole waves of
    // resident CTAs, charging ~4 % of a CTA per extra split for the table load / final select.
    {
      double best = -1.0;
      for (int s = 1; s <= 8 && s <= nprobe; s++) {
        const double waves = (double)n * s / slots;
        const double eff = waves / std::ceil(waves) - 0.04 * (s - 1);
        if (eff > best) best = eff, S = s;
      }
    }
    // positional work plan (v2 / M = 64): whole waves unsplit, only the queries of the last partial wave are split
    plan = variant == 2 && n <= 4096 && !T.splits && !T.no_plan;
    if (plan) {
      int rows = 1;
      plan_work(n, slots, nprobe, R, S, T.tail, &n_full, &s_tail, &n_items, &rows);
      S = rows;
    }
    if (T.splits) S = T.splits;
    S = std::max(1, std::min(S, nprobe));
    while (S > 1 && (long long)S * R > 8192) S--;
    P.max_np_s = (plan && n_full > 0) ? nprobe : (nprobe + S - 1) / S;
  }
  P.S = S;
  CKI(c.ws_cand.ensure((size_t)n * S * R * sizeof(u64)));
  P.cand = c.ws_cand.as<u64>();
  const size_t smem_need = variant == 3 ? scan_v3_smem_bytes_for(nprobe, P.v3_max_items, cap, threads) : scan_smem_bytes(P, ix->mode);
  if (smem_need > 227 * 1024) {
    set_err("scan needs %zu B shared memory (M=%d recall_num=%d nprobe=%d): not implemented", smem_need, M, R, nprobe);
    return GB200_EUNSUPPORTED;
  }
  const bool pre_zeroed = c.zeroed_for == n && variant == 3;
  c.zeroed_for = 0;
  if (!pre_zeroed) CK(cudaMemsetAsync(c.d_scanned, 0, sizeof(unsigned long long), c.stream));
  if (c.timed) CK(cudaEventRecord(c.ev[1], c.stream));
  const int *d_rows = nullptr;
  if (ix->mode == 1 || ix->mode == 2) {
    // per-query lookup tables: built on the side stream during the coarse stage (ivfpq_search_impl), else here
    const size_t lut_bytes = ix->mode == 2 ? 98304 : 65536;
    if (c.lut_built_n == n && c.lut_built_ip == (ip ? 1 : 0)) {
      CK(cudaStreamWaitEvent(c.stream, c.ev_join, 0));
    } else {
      CKI(c.ws_lut.ensure((size_t)n * lut_bytes));
      if (ix->mode == 2)
        CK(launch_lut_build_m64(d_xq, ix->d_pq_t, c.ws_lut.as<float>(), n, ix->p.d, ix->dsub, ip ? 1 : 0, c.stream));
      else
        CK(launch_lut_build_m32(d_xq, ix->d_pq_t, c.ws_lut.as<float>(), n, ix->p.d, ix->dsub, ip ? 1 : 0, c.stream));
      c.launches++;
    }
    c.lut_built_n = 0;
    P.lut_g = c.ws_lut.as<float>();
    if (variant == 3) {
      // control words [next_q, pad x3][claim x n][rows x n], zeroed; per-query probe tables
      CKI(c.ws_ctl.ensure((size_t)(4 + 2 * (size_t)n) * sizeof(int)));
      if (!pre_zeroed) CK(cudaMemsetAsync(c.ws_ctl.p, 0, (size_t)(4 + 2 * (size_t)n) * sizeof(int), c.stream));
      P.v3_next_q = c.ws_ctl.as<int>();
      P.v3_claim = c.ws_ctl.as<int>() + 4;
      P.v3_rows = c.ws_ctl.as<int>() + 4 + n;
      d_rows = P.v3_rows;
      CKI(c.ws_probe.ensure((size_t)n * scan_v3_probe_bytes(nprobe, P.v3_max_items)));
      P.probe_g = c.ws_probe.as<unsigned char>();
      CK(launch_probe_setup_v3(P, c.stream));
      c.launches++;
    } else {
      if (plan) {
        P.n_items = n_items;
        P.n_full = n_full;
        P.s_tail = s_tail;
      }
      CKI(c.ws_probe.ensure((size_t)(plan ? n_items : n * S) * scan_probe_bytes_host(P.max_np_s)));
      P.probe_g = c.ws_probe.as<unsigned char>();
      CK(launch_probe_setup(P, c.stream));
      c.launches++;
    }
  }
  if (c.timed) CK(cudaEventRecord(c.ev[4], c.stream));
  if (variant == 3) CK(launch_ivfpq_scan_v3(P, grid_v3, c.stream));
  else CK(launch_ivfpq_scan(P, ix->mode, c.stream));
  if (c.timed) CK(cudaEventRecord(c.ev[5], c.stream));
  if (c.timed) CK(cudaEventRecord(c.ev[2], c.stream));
  RerankParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.no_stage = T.rerank_no_stage;
  Q.cand = P.cand;
  Q.keys = d_keys;
  Q.list_off = ix->d_off;
  Q.ids = ix->d_ids;
  Q.xq = d_xq_raw;
  Q.raw = ix->d_raw;
  Q.nraw = ix->raw_n.load();
  Q.out_dist = d_out_d;
  Q.out_ids = d_out_i;
  Q.n = n;
  Q.S = S;
  Q.R = R;
  Q.k = k;
  Q.nprobe = nprobe;
  Q.raw_d = ix->p.raw_d;
  Q.xq_stride = ix->p.d;
  Q.has_rank = sp->has_rank ? 1 : 0;
  Q.is_ip = ip ? 1 : 0;
  Q.min_score = sp->min_score;
  Q.max_score = sp->max_score;
  Q.nsplit = d_rows;  // v3: rows handed out per query (clamped to S by the kernel)
  Q.n_full = P.n_items > 0 ? P.n_full : 0;
  if (c.sink) {
    Q.sink = *c.sink;
    c.sink_used = true;
  }
  CK(launch_rerank(Q, c.stream));
  if (c.timed) CK(cudaEventRecord(c.ev[3], c.stream));
  c.launches += 2;
  return GB200_OK;
}

// wait for the context's stream and publish its counters as the index's "last call" statistics
static int finish_profile(gb200_index *ix, SearchCtx &c) {
  unsigned long long sc = 0;
  CK(cudaMemcpyAsync(&sc, c.d_scanned, sizeof(sc), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  std::lock_guard<std::mutex> g(ix->stats_mu);
  ix->last_scanned = (long long)sc;
  ix->launches += c.launches;
  c.launches = 0;
  if (c.timed) {
    cudaEventElapsedTime(&ix->stage_ms[0], c.ev[0], c.ev[1]);
    cudaEventElapsedTime(&ix->stage_ms[1], c.ev[1], c.ev[2]);
    cudaEventElapsedTime(&ix->stage_ms[2], c.ev[2], c.ev[3]);
    cudaEventElapsedTime(&ix->stage_ms[3], c.ev[0], c.ev[3]);
    cudaEventElapsedTime(&ix->scan_kernel_ms, c.ev[4], c.ev[5]);
  }
  return GB200_OK;
}

static int check_search_args(gb200_index *ix, int n, const void *xq, int k, const gb200_search_params *sp,
                             const void *D, const void *I) {
  if (!ix || n < 0 || !sp || (n > 0 && (!xq || !D || !I))) return GB200_EINVAL;
  if (k <= 0) {  // reference logs a warning and returns without touching the outputs (gamma_index_ivfpq.cc:753-756)
    set_err("topK must be > 0");
    return GB200_EINVAL;
  }
  if (sp->metric != GB200_METRIC_INNER_PRODUCT && sp->metric != GB200_METRIC_L2) return GB200_EINVAL;
  return GB200_OK;
}

static int ivfpq_search_impl(gb200_index *ix, SearchCtx &c, int n, const float *xq, bool xq_on_dev, int k,
                             const gb200_search_params *sp, const gb200_range_filter *filters, int n_filters,
                             bool use_installed_filter, const int64_t *keys_h, const float *cdis_h, int nprobe_pre,
                             float *D, int64_t *I, bool out_on_dev) {
  if (n == 0) return GB200_OK;
  if (ix->flat_lists) {
    set_err("this is an IVFFLAT index: use gb200_ivfflat_search");
    return GB200_EINVAL;
  }
  if (!ix->trained) {
    set_err("IVFPQ search on an untrained index");
    return GB200_ENOTTRAINED;
  }
  const int d = ix->p.d;
  int nprobe = keys_h ? nprobe_pre : resolve_nprobe(ix, sp, ix->p.nprobe > 0 ? ix->p.nprobe : 80);
  if (nprobe <= 0) return GB200_EINVAL;
  c.timed = ix->profiling;
  const float *d_xq = xq;
  if (!xq_on_dev) {
    CKI(c.ws_xq.ensure((size_t)n * d * sizeof(float)));
    CK(cudaMemcpyAsync(c.ws_xq.p, xq, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    d_xq = c.ws_xq.as<float>();
  }
  const float *d_xq_raw = d_xq;  // the exact re-rank compares the query as given with the raw vectors
  if (ix->d_opq_At) {            // opq_->apply(n, xq) (gamma_index_ivfpq.cc:547-555)
    CKI(c.ws_xt.ensure((size_t)n * d * sizeof(float)));
    CK(launch_linear_apply(d_xq, d, n, d, ix->d_opq_At, ix->d_opq_b, d, c.ws_xt.as<float>(), c.stream));
    c.launches++;
    d_xq = c.ws_xt.as<float>();
  }
  const uint32_t *d_valid = nullptr;
  long long valid_bits = 0;
  if (use_installed_filter && ix->inst.active) {
    d_valid = ix->inst.valid.as<uint32_t>();
    valid_bits = ix->inst.bits;
  } else {
    CKI(prepare_valid(ix, c, filters, n_filters, &d_valid, &valid_bits));
  }
  CKI(c.ws_keys.ensure((size_t)n * nprobe * sizeof(int)));
  CKI(c.ws_cdis.ensure((size_t)n * nprobe * sizeof(float)));
  if (c.timed) CK(cudaEventRecord(c.ev[0], c.stream));
  c.lut_built_n = 0;
  if ((ix->mode == 1 || ix->mode == 2) && d <= 1024 && !keys_h && !ix->tune.lut_inline) {
    // K2a depends on the queries only: fork it onto the side stream so it overlaps the coarse quantiser
    const int ipm = sp->metric == GB200_METRIC_INNER_PRODUCT ? 1 : 0;
    CKI(c.ws_lut.ensure((size_t)n * (ix->mode == 2 ? 98304 : 65536)));
    CK(cudaEventRecord(c.ev_fork, c.stream));
    CK(cudaStreamWaitEvent(c.stream2, c.ev_fork, 0));
    if (ix->mode == 2)
      CK(launch_lut_build_m64(d_xq, ix->d_pq_t, c.ws_lut.as<float>(), n, d, ix->dsub, ipm, c.stream2));
    else
      CK(launch_lut_build_m32(d_xq, ix->d_pq_t, c.ws_lut.as<float>(), n, d, ix->dsub, ipm, c.stream2));
    CK(cudaEventRecord(c.ev_join, c.stream2));
    c.launches++;
    c.lut_built_n = n;
    c.lut_built_ip = ipm;
  }
  if (keys_h) {
    std::vector<int> k32((size_t)n * nprobe);
    for (size_t i = 0; i < k32.size(); i++) k32[i] = (keys_h[i] < 0 || keys_h[i] >= ix->p.nlist) ? -1 : (int)keys_h[i];
    CK(cudaMemcpyAsync(c.ws_keys.p, k32.data(), k32.size() * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    CK(cudaMemcpyAsync(c.ws_cdis.p, cdis_h, (size_t)n * nprobe * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    CK(cudaStreamSynchronize(c.stream));
  } else {
    c.zero_req = 0;
    if (ix->mode == 1) {  // the persistent scan's control words are zeroed by the coarse stage's first launch
      CKI(c.ws_ctl.ensure((size_t)(4 + 2 * (size_t)n) * sizeof(int)));
      c.zero_req = n;
    }
    CKI(coarse_dev(ix, c, n, d_xq, nprobe, c.ws_keys.as<int>(), c.ws_cdis.as<float>()));
    c.zero_req = 0;
  }
  float *d_D = D;
  long long *d_I = reinterpret_cast<long long *>(I);
  if (!out_on_dev) {
    CKI(c.ws_out_d.ensure((size_t)n * k * sizeof(float)));
    CKI(c.ws_out_i.ensure((size_t)n * k * sizeof(long long)));
    d_D = c.ws_out_d.as<float>();
    d_I = c.ws_out_i.as<long long>();
  }
  CKI(scan_rerank_dev(ix, c, n, d_xq, d_xq_raw, k, sp, nprobe, c.ws_keys.as<int>(), c.ws_cdis.as<float>(), d_valid, valid_bits, d_D,
                      d_I));
  if (!out_on_dev) {
    CK(cudaMemcpyAsync(D, d_D, (size_t)n * k * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaMemcpyAsync(I, d_I, (size_t)n * k * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    CKI(finish_profile(ix, c));
  }
  return GB200_OK;
}

// order the context's stream after the caller's stream (device-resident entry points), and back
static int join_user_stream(SearchCtx &c, cudaStream_t cs) {
  CK(cudaEventRecord(c.ev_user, cs));
  CK(cudaStreamWaitEvent(c.stream, c.ev_user, 0));
  return GB200_OK;
}
static int release_to_user_stream(SearchCtx &c, cudaStream_t cs) {
  CK(cudaEventRecord(c.ev_user, c.stream));
  CK(cudaStreamWaitEvent(cs, c.ev_user, 0));
  return GB200_OK;
}

// One device batch for a group of callers: their queries are copied into one workspace, one search runs, each caller
// gets its rows back.  A query's result does not depend on the batch it travels in (every stage works per query row).
static int run_merged(gb200_index *ix, const std::vector<gb200_index::PendingSearch *> &grp) {
  SearchScope s(ix);
  if (!s.c) return GB200_ECUDA;
  SearchCtx &c = *s.c;
  const int d = ix->p.d, k = grp[0]->k;
  int n = 0;
  for (auto *r : grp) n += r->n;
  CKI(c.ws_xq.ensure((size_t)n * d * sizeof(float)));
  CKI(c.ws_out_d.ensure((size_t)n * k * sizeof(float)));
  CKI(c.ws_out_i.ensure((size_t)n * k * sizeof(long long)));
  const size_t out_bytes = (size_t)n * k * (sizeof(float) + sizeof(long long));
  if (c.h_stage_cap < out_bytes) {
    if (c.h_stage) CK(cudaFreeHost(c.h_stage));
    c.h_stage = nullptr, c.h_stage_cap = 0;
    CK(cudaMallocHost(&c.h_stage, out_bytes * 2));
    c.h_stage_cap = out_bytes * 2;
  }
  size_t row = 0;
  for (auto *r : grp) {
    CK(cudaMemcpyAsync(c.ws_xq.as<float>() + row * d, r->xq, (size_t)r->n * d * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    row += r->n;
  }
  CKI(ivfpq_search_impl(ix, c, n, c.ws_xq.as<float>(), true, k, &grp[0]->sp, nullptr, 0, false, nullptr, nullptr, 0,
                        c.ws_out_d.as<float>(), reinterpret_cast<int64_t *>(c.ws_out_i.p), true));
  long long *h_I = static_cast<long long *>(c.h_stage);
  float *h_D = reinterpret_cast<float *>(h_I + (size_t)n * k);
  CK(cudaMemcpyAsync(h_I, c.ws_out_i.p, (size_t)n * k * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaMemcpyAsync(h_D, c.ws_out_d.p, (size_t)n * k * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
  CKI(finish_profile(ix, c));
  row = 0;
  for (auto *r : grp) {
    memcpy(r->I, h_I + row * k, (size_t)r->n * k * sizeof(long long));
    memcpy(r->D, h_D + row * k, (size_t)r->n * k * sizeof(float));
    row += r->n;
  }
  return GB200_OK;
}

// Group commit for concurrent callers: the policy lives in coalesce.h (host-only, stress-tested under ThreadSanitizer by
// tests/test_coalesce_cpu.py); here only what a batch of requests runs.
static int search_coalesced(gb200_index *ix, int n, const float *xq, int k, const gb200_search_params *sp, float *D, int64_t *I) {
  gb200_index::PendingSearch me;
  me.n = n, me.k = k, me.xq = xq, me.D = D, me.I = I, me.sp = *sp;
  gb::CoalescePolicy pol;
  pol.slots = ix->tune.coalesce;
  pol.max_queries = ix->tune.coalesce_max;
  pol.wait_us = ix->tune.coalesce_wait_us;
  pol.balance = ix->tune.coalesce_balance;
  std::string err;
  const int rc = ix->coalescer.submit(
      me, pol,
      [&](const std::vector<gb200_index::PendingSearch *> &grp, std::string &e) {
        int r;
        if (grp.size() == 1) {  // nobody to take along: the plain path, results straight into the caller's buffers
          SearchScope s(ix);
          r = !s.c ? GB200_ECUDA
                   : ivfpq_search_impl(ix, *s.c, n, xq, false, k, sp, nullptr, 0, false, nullptr, nullptr, 0, D, I, false);
        } else {
          r = run_merged(ix, grp);
        }
        if (r != GB200_OK) e = g_err;
        return r;
      },
      &err);
  if (rc != GB200_OK && !err.empty()) set_err("%s", err.c_str());  // this request travelled in another caller's batch
  return rc;
}

int gb200_ivfpq_search(gb200_index *ix, int n, const float *xq, int k, const gb200_search_params *sp,
                       const gb200_range_filter *filters, int n_filters, float *D, int64_t *I) {
  CKI(check_search_args(ix, n, xq, k, sp, D, I));
  if (ix->kind != 0) return GB200_EINVAL;
  CKI(use_device(ix));
  if (ix->tune.coalesce > 0 && n > 0 && n_filters == 0 && 2 * n <= ix->tune.coalesce_max && !ix->profiling && !ix->flat_lists &&
      ix->trained)
    return search_coalesced(ix, n, xq, k, sp, D, I);
  SearchScope s(ix);
  if (!s.c) return GB200_ECUDA;
  return ivfpq_search_impl(ix, *s.c, n, xq, false, k, sp, filters, n_filters, false, nullptr, nullptr, 0, D, I, false);
}

int gb200_ivfpq_search_preassigned(gb200_index *ix, int n, const float *xq, int k, const gb200_search_params *sp,
                                   const gb200_range_filter *filters, int n_filters, const int64_t *keys,
                                   const float *coarse_dis, int nprobe, float *D, int64_t *I) {
  CKI(check_search_args(ix, n, xq, k, sp, D, I));
  if (ix->kind != 0 || !keys || !coarse_dis || nprobe <= 0) return GB200_EINVAL;
  CKI(use_device(ix));
  SearchScope s(ix);
  if (!s.c) return GB200_ECUDA;
  return ivfpq_search_impl(ix, *s.c, n, xq, false, k, sp, filters, n_filters, false, keys, coarse_dis, nprobe, D, I, false);
}

int gb200_ivfpq_search_dev(gb200_index *ix, int n, const float *xq_dev, int k, const gb200_search_params *sp,
                           float *D_dev, int64_t *I_dev, void *stream) {
  return gb_ivfpq_search_dev_sink(ix, n, xq_dev, k, sp, D_dev, I_dev, stream, nullptr, nullptr);
}

int gb200_ivfpq_coarse(gb200_index *ix, int n, const float *xq, int nprobe, float *coarse_dis, int64_t *keys) {
  if (!ix || ix->kind != 0 || n < 0 || nprobe <= 0 || nprobe > ix->p.nlist || !xq || !coarse_dis || !keys)
    return GB200_EINVAL;
  if (!ix->trained) return GB200_ENOTTRAINED;
  if (n == 0) return GB200_OK;
  CKI(use_device(ix));
  SearchScope s(ix);
  if (!s.c) return GB200_ECUDA;
  SearchCtx &c = *s.c;
  const int d = ix->p.d;
  CKI(c.ws_xq.ensure((size_t)n * d * sizeof(float)));
  CK(cudaMemcpyAsync(c.ws_xq.p, xq, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, c.stream));
  const float *d_q = c.ws_xq.as<float>();
  if (ix->d_opq_At) {
    CKI(c.ws_xt.ensure((size_t)n * d * sizeof(float)));
    CK(launch_linear_apply(d_q, d, n, d, ix->d_opq_At, ix->d_opq_b, d, c.ws_xt.as<float>(), c.stream));
    d_q = c.ws_xt.as<float>();
  }
  CKI(c.ws_keys.ensure((size_t)n * nprobe * sizeof(int)));
  CKI(c.ws_cdis.ensure((size_t)n * nprobe * sizeof(float)));
  CKI(coarse_dev(ix, c, n, d_q, nprobe, c.ws_keys.as<int>(), c.ws_cdis.as<float>()));
  std::vector<int> k32((size_t)n * nprobe);
  CK(cudaMemcpyAsync(k32.data(), c.ws_keys.p, k32.size() * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaMemcpyAsync(coarse_dis, c.ws_cdis.p, (size_t)n * nprobe * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  for (size_t i = 0; i < k32.size(); i++) keys[i] = k32[i];
  ix->launches += c.launches;
  c.launches = 0;
  return GB200_OK;
}

// ---- encode (stage 1 of Add) ---------------------------------------------------------------------------
// rows already on the device (x_stride floats per row, x_stride <= d); outputs to host
static int encode_dev(gb200_index *ix, SearchCtx &c, int64_t n, const float *d_x, int x_stride, int32_t *list_no,
                      uint8_t *codes) {
  const int d = ix->p.d, M = ix->p.nsubvector;
  const int64_t CH = 1 << 17;
  for (int64_t s0 = 0; s0 < n; s0 += CH) {
    const int m = (int)std::min<int64_t>(CH, n - s0);
    const float *rows = d_x + (size_t)s0 * x_stride;
    if (ix->d_opq_At) {  // opq_->apply (gamma_index_ivfpq.cc:448-450); reads x_stride columns, zero beyond
      CKI(c.ws_xt.ensure((size_t)m * d * sizeof(float)));
      CK(launch_linear_apply(rows, x_stride, m, d, ix->d_opq_At, ix->d_opq_b, d, c.ws_xt.as<float>(), c.stream));
      c.launches++;
      rows = c.ws_xt.as<float>();
    } else if (x_stride != d) {  // ConvertVectorDim: zero-pad to d for the distance producer
      CKI(c.ws_xq.ensure((size_t)m * d * sizeof(float)));
      CK(cudaMemsetAsync(c.ws_xq.p, 0, (size_t)m * d * sizeof(float), c.stream));
      CK(cudaMemcpy2DAsync(c.ws_xq.p, (size_t)d * sizeof(float), rows, (size_t)x_stride * sizeof(float),
                           (size_t)x_stride * sizeof(float), m, cudaMemcpyDeviceToDevice, c.stream));
      rows = c.ws_xq.as<float>();
    }
    const int stride = (ix->d_opq_At || x_stride != d) ? d : x_stride;
    CKI(c.ws_keys.ensure((size_t)m * sizeof(int)));
    CKI(c.ws_cdis.ensure((size_t)m * sizeof(float)));
    CKI(coarse_dev(ix, c, m, rows, 1, c.ws_keys.as<int>(), c.ws_cdis.as<float>()));
    CKI(c.ws_cand.ensure((size_t)m * M));
    CK(launch_pq_encode(rows, stride, c.ws_keys.as<int>(), ix->d_cent, ix->d_pq, m, d, M, ix->dsub, 1,
                        c.ws_cand.as<uint8_t>(), c.stream));
    c.launches++;
    CK(cudaMemcpyAsync(list_no + s0, c.ws_keys.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaMemcpyAsync(codes + (size_t)s0 * M, c.ws_cand.p, (size_t)m * M, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
  }
  ix->launches += c.launches;
  c.launches = 0;
  return GB200_OK;
}

int gb200_ivfpq_encode(gb200_index *ix, int64_t n, const float *x, int x_dim, int32_t *list_no, uint8_t *codes) {
  if (!ix || ix->kind != 0 || n < 0 || x_dim <= 0 || x_dim > ix->p.d || (n > 0 && (!x || !list_no || !codes)))
    return GB200_EINVAL;
  if (!ix->trained) return GB200_ENOTTRAINED;
  if (n == 0) return GB200_OK;
  CKI(use_device(ix));
  SearchScope s(ix);
  if (!s.c) return GB200_ECUDA;
  SearchCtx &c = *s.c;
  const int64_t CH = 1 << 17;
  for (int64_t s0 = 0; s0 < n; s0 += CH) {
    const int64_t m = std::min<int64_t>(CH, n - s0);
    CKI(c.ws_flat.ensure((size_t)m * x_dim * sizeof(float)));
    CK(cudaMemcpyAsync(c.ws_flat.p, x + (size_t)s0 * x_dim, (size_t)m * x_dim * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    CKI(encode_dev(ix, c, m, c.ws_flat.as<float>(), x_dim, list_no + s0, codes + (size_t)s0 * ix->p.nsubvector));
  }
  return GB200_OK;
}

int gb200_ivfpq_add_stored(gb200_index *ix, int64_t first_vid, int64_t n, int32_t *list_no, uint8_t *codes) {
  if (!ix || ix->kind != 0 || first_vid < 0 || n < 0) return GB200_EINVAL;
  if (!ix->trained) return GB200_ENOTTRAINED;
  if (n == 0) return GB200_OK;
  if (first_vid + n > ix->raw_n.load()) {
    set_err("add_stored: rows %lld .. %lld are not in the raw store (%lld rows)", (long long)first_vid,
            (long long)(first_vid + n - 1), (long long)ix->raw_n.load());
    return GB200_EINVAL;
  }
  const int M = ix->p.nsubvector, rd = ix->p.raw_d;
  std::vector<int32_t> ln_own;
  std::vector<uint8_t> cd_own;
  if (!list_no) {
    ln_own.resize((size_t)n);
    list_no = ln_own.data();
  }
  if (!codes) {
    cd_own.resize((size_t)n * M);
    codes = cd_own.data();
  }
  {
    CKI(use_device(ix));
    SearchScope s(ix);  // shared hold: the raw store cannot be replaced while the rows are read
    if (!s.c) return GB200_ECUDA;
    CKI(encode_dev(ix, *s.c, n, ix->d_raw + (size_t)first_vid * rd, rd, list_no, codes));
  }
  std::vector<int64_t> vids((size_t)n);
  for (int64_t i = 0; i < n; i++) vids[i] = first_vid + i;
  return gb200_ivfpq_append(ix, n, list_no, vids.data(), codes);
}

int gb200_ivfpq_add_raw(gb200_index *ix, int64_t first_vid, int64_t n, const float *x, int32_t *list_no, uint8_t *codes) {
  if (!ix || ix->kind != 0 || first_vid < 0 || n < 0 || (n > 0 && !x)) return GB200_EINVAL;
  if (!ix->trained) return GB200_ENOTTRAINED;
  if (n == 0) return GB200_OK;
  // the raw rows first (re-rank and flat read them; the encode reads them from the device store)
  CKI(gb200_upload_raw(ix, first_vid, n, x));
  return gb200_ivfpq_add_stored(ix, first_vid, n, list_no, codes);
}

// ---- IVFFLAT ------------------------------------------------------------------------------------------
int gb200_ivfflat_create(int device, int d, int nlist, int metric, int nprobe, gb200_index **out) {
  if (!out || d <= 0 || nlist <= 0) return GB200_EINVAL;
  if (d % 4) {
    set_err("IVFFLAT: d = %d is not a multiple of 4 (not implemented)", d);
    return GB200_EUNSUPPORTED;
  }
  gb200_ivfpq_params p;
  memset(&p, 0, sizeof(p));
  p.device = device;
  p.d = d;
  p.raw_d = d;
  p.nlist = nlist;
  p.nsubvector = 4;  // dummy codes: the posting machinery (blocks, growth, publication, compaction) is shared
  p.nbits = 8;
  p.metric = metric;
  p.nprobe = nprobe;
  p.store_raw = 1;
  int rc = gb200_ivfpq_create(&p, out);
  if (rc == GB200_OK) (*out)->flat_lists = true;
  return rc;
}

int gb200_ivfflat_set_quantizer(gb200_index *ix, const float *coarse) {
  if (!ix || ix->kind != 0 || !ix->flat_lists || !coarse) return GB200_EINVAL;
  std::vector<float> pq((size_t)256 * ix->p.d, 0.f);
  return gb200_ivfpq_set_quantizers(ix, coarse, pq.data());
}

int gb200_ivfflat_append(gb200_index *ix, int64_t n, const int32_t *list_no, const int64_t *vids) {
  if (!ix || ix->kind != 0 || !ix->flat_lists || n < 0) return GB200_EINVAL;
  std::vector<uint8_t> codes((size_t)n * ix->p.nsubvector, 0);
  return gb200_ivfpq_append(ix, n, list_no, vids, codes.data());
}

int gb200_ivfflat_add_raw(gb200_index *ix, int64_t first_vid, int64_t n, const float *x, int32_t *list_no) {
  if (!ix || ix->kind != 0 || !ix->flat_lists || first_vid < 0 || n < 0 || (n > 0 && !x)) return GB200_EINVAL;
  if (!ix->trained) return GB200_ENOTTRAINED;
  if (n == 0) return GB200_OK;
  CKI(gb200_upload_raw(ix, first_vid, n, x));
  std::vector<int32_t> ln_own;
  if (!list_no) {
    ln_own.resize((size_t)n);
    list_no = ln_own.data();
  }
  {
    CKI(use_device(ix));
    SearchScope s(ix);
    if (!s.c) return GB200_ECUDA;
    SearchCtx &c = *s.c;
    const int d = ix->p.d;
    const int64_t CH = 1 << 17;
    for (int64_t s0 = 0; s0 < n; s0 += CH) {  // quantizer->assign = the coarse stage with nprobe = 1
      const int m = (int)std::min<int64_t>(CH, n - s0);
      CKI(c.ws_keys.ensure((size_t)m * sizeof(int)));
      CKI(c.ws_cdis.ensure((size_t)m * sizeof(float)));
      CKI(coarse_dev(ix, c, m, ix->d_raw + (size_t)(first_vid + s0) * d, 1, c.ws_keys.as<int>(), c.ws_cdis.as<float>()));
      CK(cudaMemcpyAsync(list_no + s0, c.ws_keys.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
      CK(cudaStreamSynchronize(c.stream));
    }
    ix->launches += c.launches;
    c.launches = 0;
  }
  std::vector<int64_t> vids((size_t)n);
  for (int64_t i = 0; i < n; i++) vids[i] = first_vid + i;
  return gb200_ivfflat_append(ix, n, list_no, vids.data());
}

int gb200_ivfflat_search(gb200_index *ix, int n, const float *xq, int k, const gb200_search_params *sp,
                         const gb200_range_filter *filters, int n_filters, float *D, int64_t *I) {
  CKI(check_search_args(ix, n, xq, k, sp, D, I));
  if (ix->kind != 0 || !ix->flat_lists) return GB200_EINVAL;
  if (n == 0) return GB200_OK;
  if (!ix->trained) {
    set_err("IVFFLAT search on an untrained index");
    return GB200_ENOTTRAINED;
  }
  if (k > 2048) {
    set_err("IVFFLAT k=%d > 2048 not implemented", k);
    return GB200_EUNSUPPORTED;
  }
  CKI(use_device(ix));
  SearchScope s(ix);
  if (!s.c) return GB200_ECUDA;
  SearchCtx &c = *s.c;
  const int d = ix->p.d;
  const int nprobe = resolve_nprobe(ix, sp, ix->p.nprobe > 0 ? ix->p.nprobe : 80);
  const bool ip = sp->metric == GB200_METRIC_INNER_PRODUCT;
  c.timed = ix->profiling;
  CKI(c.ws_xq.ensure((size_t)n * d * sizeof(float)));
  CK(cudaMemcpyAsync(c.ws_xq.p, xq, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, c.stream));
  const float *d_xq = c.ws_xq.as<float>();
  const uint32_t *d_valid = nullptr;
  long long valid_bits = 0;
  CKI(prepare_valid(ix, c, filters, n_filters, &d_valid, &valid_bits));
  CKI(c.ws_keys.ensure((size_t)n * nprobe * sizeof(int)));
  CKI(c.ws_cdis.ensure((size_t)n * nprobe * sizeof(float)));
  if (c.timed) CK(cudaEventRecord(c.ev[0], c.stream));
  CKI(coarse_dev(ix, c, n, d_xq, nprobe, c.ws_keys.as<int>(), c.ws_cdis.as<float>()));
  if (c.timed) CK(cudaEventRecord(c.ev[1], c.stream));
  // splits: fill the machine when the batch is small (one CTA per (query, split), probes dealt round-robin)
  int S = 1;
  const int slots = ix->num_sms * 4;
  while (S < 8 && S * 2 <= nprobe && (long long)n * S * 2 <= slots) S *= 2;
  IvfFlatParams F;
  memset(&F, 0, sizeof(F));
  F.xq = d_xq;
  F.keys = c.ws_keys.as<int>();
  F.list_off = ix->d_off;
  F.list_len = ix->d_len;
  F.ids = ix->d_ids;
  F.raw = ix->d_raw;
  F.nraw = ix->raw_n.load();
  F.valid = d_valid;
  F.valid_bits = valid_bits;
  F.scanned = c.d_scanned;
  F.n = n, F.d = d, F.nlist = ix->p.nlist, F.nprobe = nprobe, F.S = S, F.R = k;
  F.cap = ivfflat_buffer_cap(k);
  F.is_ip = ip ? 1 : 0;
  F.min_score = sp->min_score, F.max_score = sp->max_score;
  CKI(c.ws_cand.ensure((size_t)n * S * k * sizeof(u64)));
  F.cand = c.ws_cand.as<u64>();
  CK(cudaMemsetAsync(c.d_scanned, 0, sizeof(unsigned long long), c.stream));
  if (c.timed) CK(cudaEventRecord(c.ev[4], c.stream));
  CK(launch_ivfflat_scan(F, c.stream));
  if (c.timed) {
    CK(cudaEventRecord(c.ev[5], c.stream));
    CK(cudaEventRecord(c.ev[2], c.stream));
  }
  CKI(c.ws_out_d.ensure((size_t)n * k * sizeof(float)));
  CKI(c.ws_out_i.ensure((size_t)n * k * sizeof(long long)));
  RerankParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.cand = F.cand;
  Q.keys = F.keys;
  Q.list_off = ix->d_off;
  Q.ids = ix->d_ids;
  Q.xq = d_xq;
  Q.raw = ix->d_raw;
  Q.nraw = F.nraw;
  Q.out_dist = c.ws_out_d.as<float>();
  Q.out_ids = c.ws_out_i.as<long long>();
  Q.n = n, Q.S = S, Q.R = k, Q.k = k, Q.nprobe = nprobe, Q.raw_d = d, Q.xq_stride = d;
  Q.has_rank = 0;  // the scan's distances are exact already: merge the rows, resolve vids, first k
  Q.is_ip = ip ? 1 : 0;
  Q.min_score = sp->min_score, Q.max_score = sp->max_score;
  CK(launch_rerank(Q, c.stream));
  if (c.timed) CK(cudaEventRecord(c.ev[3], c.stream));
  c.launches += 2;
  CK(cudaMemcpyAsync(D, Q.out_dist, (size_t)n * k * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaMemcpyAsync(I, Q.out_ids, (size_t)n * k * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
  return finish_profile(ix, c);
}

// ---- flat ----------------------------------------------------------------------------------------
// x - tf32(x) and |x|^2 of the raw rows [aux_n, rows), kept next to the raw store once a batched flat search needs them.
// Index-level state shared by all searches: built under aux_mu on the caller's stream, which is then drained so that
// every other context may read the rows.
static int ensure_raw_aux(gb200_index *ix, SearchCtx &c, long long rows) {
  if (ix->aux_n.load() >= rows && ix->aux_cap >= ix->raw_cap) return GB200_OK;
  std::lock_guard<std::mutex> a(ix->aux_mu);
  const int d = ix->p.raw_d;
  if (ix->aux_cap < ix->raw_cap) {
    // other searches may be reading the old companions (they hold data_mu shared like this call): allocate the new ones,
    // copy what exists, and retire the old ones only after a device-wide drain
    float *ns = nullptr, *nn = nullptr;
    CK(cudaMalloc(&ns, (size_t)ix->raw_cap * d * sizeof(float)));
    CK(cudaMalloc(&nn, (size_t)ix->raw_cap * sizeof(float)));
    const long long have = ix->aux_n.load();
    if (have > 0) {
      CK(cudaMemcpyAsync(ns, ix->d_raw_small, (size_t)have * d * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
      CK(cudaMemcpyAsync(nn, ix->d_raw_norm, (size_t)have * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
    }
    CK(cudaDeviceSynchronize());
    if (ix->d_raw_small) cudaFree(ix->d_raw_small);
    if (ix->d_raw_norm) cudaFree(ix->d_raw_norm);
    ix->d_raw_small = ns;
    ix->d_raw_norm = nn;
    ix->aux_cap = ix->raw_cap;
  }
  const long long r0 = ix->aux_n.load();
  if (r0 < rows) {
    const long long m = rows - r0;
    CK(launch_tf32_residual(ix->d_raw + (size_t)r0 * d, ix->d_raw_small + (size_t)r0 * d, (size_t)m * d, c.stream));
    for (long long s0 = 0; s0 < m; s0 += (1 << 30) / 8) {  // row_norms takes int rows
      int nr = (int)std::min<long long>((1 << 30) / 8, m - s0);
      CK(launch_row_norms(ix->d_raw + (size_t)(r0 + s0) * d, nr, d, ix->d_raw_norm + r0 + s0, c.stream));
    }
    c.launches += 2;
    CK(cudaStreamSynchronize(c.stream));
    ix->aux_n = rows;
  }
  return GB200_OK;
}

// batched FLAT: tcgen05 3xTF32 GEMM per database chunk -> running candidate select -> exact re-score (flat_tc.cu)
static int flat_tc_dev(gb200_index *ix, SearchCtx &c, int n, const float *d_xq, int k, const gb200_search_params *sp,
                       const uint32_t *d_valid, long long N, float *d_D, long long *d_I) {
  const int d = ix->p.raw_d;
  const bool ip = sp->metric == GB200_METRIC_INNER_PRODUCT;
  CKI(ensure_raw_aux(ix, c, N));
  CKI(c.ws_xs.ensure((size_t)n * d * sizeof(float)));
  CKI(c.ws_xn.ensure((size_t)n * sizeof(float)));
  CK(launch_rows_prep(d_xq, n, d, c.ws_xn.as<float>(), c.ws_xs.as<float>(), c.stream));
  c.launches += 1;
  long long nc_max = ((1LL << 26) / n) & ~127LL;  // distance tile <= 256 MB
  if (nc_max < 1024) nc_max = 1024;
  if (ix->tune.flat_chunk_rows > 0) nc_max = std::max(1024LL, ix->tune.flat_chunk_rows & ~127LL);  // tests: force many chunks
  if (nc_max > N) nc_max = (N + 127) & ~127LL;
  CKI(c.ws_dist.ensure((size_t)n * nc_max * sizeof(float)));
  const int Kp = flat_tc_candidates(k);
  CKI(c.ws_fstate.ensure((size_t)n * Kp * sizeof(u64)));
  // candidates are taken with a slightly widened window; the exact window is applied after the re-score
  auto widen = [](float v, float sign) {
    if (!(fabsf(v) < 1e30f)) return v;
    return v + sign * (1e-3f * fabsf(v) + 1e-6f);
  };
  const float lo = widen(sp->min_score, -1.f), hi = widen(sp->max_score, 1.f);
  int first = 1;
  for (long long c0 = 0; c0 < N; c0 += nc_max) {
    const int nc = (int)std::min<long long>(nc_max, N - c0);
    CK(launch_tc_gemm(d_xq, c.ws_xs.as<float>(), c.ws_xn.as<float>(), ix->d_raw + (size_t)c0 * d,
                      ix->d_raw_small + (size_t)c0 * d, ix->d_raw_norm + c0, n, nc, d, c.ws_dist.as<float>(),
                      (int)nc_max, ip ? 0 : 1, nullptr, 0, c.stream));
    CK(launch_flat_chunk_select(c.ws_dist.as<float>(), (int)nc_max, nc, c0, d_valid, lo, hi, Kp, first,
                                c.ws_fstate.as<u64>(), n, ip ? 1 : 0, c.stream));
    c.launches += 2;
    first = 0;
  }
  CK(launch_flat_rescore(c.ws_fstate.as<u64>(), Kp, d_xq, ix->d_raw, n, d, sp->min_score, sp->max_score, k, ip ? 1 : 0,
                         d_D, d_I, c.stream));
  c.launches += 1;
  return GB200_OK;
}

static int flat_impl(gb200_index *ix, SearchCtx &c, int n, const float *xq, bool xq_on_dev, int k,
                     const gb200_search_params *sp, const gb200_range_filter *filters, int n_filters,
                     bool use_installed_filter, float *D, int64_t *I, bool out_on_dev) {
  if (n == 0) return GB200_OK;
  const int d = ix->p.raw_d;
  if (k > 2048) {
    set_err("flat k=%d > 2048 not implemented", k);
    return GB200_EUNSUPPORTED;
  }
  const long long N = ix->raw_n.load();  // the rows this search scans (rows published later are not seen)
  c.timed = ix->profiling;
  const float *d_xq = xq;
  if (!xq_on_dev) {
    CKI(c.ws_xq.ensure((size_t)n * d * sizeof(float)));
    CK(cudaMemcpyAsync(c.ws_xq.p, xq, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    d_xq = c.ws_xq.as<float>();
  }
  const uint32_t *d_valid = nullptr;
  long long valid_bits = 0;
  if (use_installed_filter && ix->inst.active) {
    d_valid = ix->inst.valid.as<uint32_t>();
    valid_bits = ix->inst.bits;
  } else {
    CKI(prepare_valid(ix, c, filters, n_filters, &d_valid, &valid_bits));
  }
  if (d_valid && valid_bits < N) {  // writers grow the bitmaps before they publish rows; never read past one
    set_err("validity bitmap covers %lld docs, store has %lld", valid_bits, N);
    return GB200_EINVAL;
  }
  float *d_D = D;
  long long *d_I = reinterpret_cast<long long *>(I);
  if (!out_on_dev) {
    CKI(c.ws_out_d.ensure((size_t)n * k * sizeof(float)));
    CKI(c.ws_out_i.ensure((size_t)n * k * sizeof(long long)));
    d_D = c.ws_out_d.as<float>();
    d_I = c.ws_out_i.as<long long>();
  }
  if (c.timed) {
    CK(cudaEventRecord(c.ev[0], c.stream));
    CK(cudaEventRecord(c.ev[1], c.stream));
    CK(cudaEventRecord(c.ev[4], c.stream));
    CK(cudaEventRecord(c.ev[5], c.stream));
  }
  // batches go through the tensor cores (GB200_FLAT=exact forces the per-query exact scan)
  const bool force_exact = ix->tune.flat_mode == 1, force_tc = ix->tune.flat_mode == 2;
  if (!force_exact && (n >= 16 || force_tc) && (d % 4 == 0) && k <= 1024) {
    CKI(flat_tc_dev(ix, c, n, d_xq, k, sp, d_valid, N, d_D, d_I));
  } else {
    FlatParams F;
    F.xq = d_xq;
    F.raw = ix->d_raw;
    F.valid = d_valid;
    F.N = N;
    F.out_dist = d_D;
    F.out_ids = d_I;
    F.n = n;
    F.d = d;
    F.k = k;
    F.is_ip = sp->metric == GB200_METRIC_INNER_PRODUCT ? 1 : 0;
    F.min_score = sp->min_score;
    F.max_score = sp->max_score;
    F.nsplit = flat_exact_splits(N, n);
    while (F.nsplit > 1 && (long long)F.nsplit * k > 8192) F.nsplit--;
    CKI(c.ws_flat.ensure((size_t)n * F.nsplit * k * sizeof(u64)));
    F.scratch = c.ws_flat.as<u64>();
    CK(launch_flat_exact(F, c.stream));
    c.launches += 2;
  }
  if (c.timed) {
    CK(cudaEventRecord(c.ev[2], c.stream));
    CK(cudaEventRecord(c.ev[3], c.stream));
  }
  if (!out_on_dev) {
    CK(cudaMemcpyAsync(D, d_D, (size_t)n * k * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaMemcpyAsync(I, d_I, (size_t)n * k * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    CKI(finish_profile(ix, c));
  }
  return GB200_OK;
}

int gb200_flat_search(gb200_index *ix, int n, const float *xq, int k, const gb200_search_params *sp,
                      const gb200_range_filter *filters, int n_filters, float *D, int64_t *I) {
  CKI(check_search_args(ix, n, xq, k, sp, D, I));
  CKI(use_device(ix));
  SearchScope s(ix);
  if (!s.c) return GB200_ECUDA;
  if (!ix->d_raw || ix->raw_n.load() == 0) {  // empty store: all slots unfilled
    for (long long i = 0; i < (long long)n * k; i++) {
      D[i] = sp->metric == GB200_METRIC_INNER_PRODUCT ? -FLT_MAX : FLT_MAX;
      I[i] = -1;
    }
    return GB200_OK;
  }
  return flat_impl(ix, *s.c, n, xq, false, k, sp, filters, n_filters, false, D, I, false);
}

int gb200_flat_search_dev(gb200_index *ix, int n, const float *xq_dev, int k, const gb200_search_params *sp,
                          float *D_dev, int64_t *I_dev, void *stream) {
  CKI(check_search_args(ix, n, xq_dev, k, sp, D_dev, I_dev));
  CKI(use_device(ix));
  SearchScope s(ix);
  if (!s.c) return GB200_ECUDA;
  if (!ix->d_raw || ix->raw_n.load() == 0) return GB200_EINVAL;
  cudaStream_t cs = (cudaStream_t)stream;
  CKI(join_user_stream(*s.c, cs));
  int rc = flat_impl(ix, *s.c, n, xq_dev, true, k, sp, nullptr, 0, true, D_dev, I_dev, true);
  if (rc == GB200_OK) CKI(release_to_user_stream(*s.c, cs));
  return rc;
}

// ---- test hook -------------------------------------------------------------------------------------
int gb200_debug_select(int device, const uint64_t *keys, int n, int R, int cap, int batch, int threads,
                       uint64_t *out, int *out_n) {
  if (!keys || !out || !out_n || n <= 0 || R <= 0 || cap < 512 || batch <= 0 || batch > cap - R ||
      cap > 16 * threads || threads % 32 != 0 || threads > 1024)
    return GB200_EINVAL;
  CK(cudaSetDevice(device));
  u64 *d_keys = nullptr, *d_out = nullptr;
  int *d_n = nullptr;
  CK(cudaMalloc(&d_keys, (size_t)n * 8));
  CK(cudaMalloc(&d_out, (size_t)R * 8));
  CK(cudaMalloc(&d_n, 4));
  CK(cudaMemcpy(d_keys, keys, (size_t)n * 8, cudaMemcpyHostToDevice));
  CK(launch_select_selftest(d_keys, n, R, cap, batch, threads, d_out, d_n, 0));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out_n, d_n, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out, d_out, (size_t)(*out_n) * 8, cudaMemcpyDeviceToHost));
  cudaFree(d_keys);
  cudaFree(d_out);
  cudaFree(d_n);
  return GB200_OK;
}

// ---- accounting ---------------------------------------------------------------------------------
int64_t gb200_mem_bytes(gb200_index *ix) {
  if (!ix) return 0;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  int64_t b = 0;
  if (ix->kind == 0) {
    b += (int64_t)ix->p.nlist * ix->p.d * 4 + (int64_t)ix->p.nlist * 4 + 2LL * ix->p.nsubvector * 256 * ix->dsub * 4;
    b += ix->pool_cap * (ix->p.nsubvector + 8) + (int64_t)ix->p.nlist * 20;
  }
  b += ix->raw_cap * ix->p.raw_d * 4;
  b += ix->live_words * 4;
  return b;
}
int64_t gb200_last_scanned_postings(gb200_index *ix) { return ix ? ix->last_scanned : 0; }
int64_t gb200_launch_count(gb200_index *ix) { return ix ? ix->launches.load() : 0; }
int gb200_last_stage_ms(gb200_index *ix, float *out4) {
  if (!ix || !out4) return GB200_EINVAL;
  std::lock_guard<std::mutex> g(ix->stats_mu);
  for (int i = 0; i < 4; i++) out4[i] = ix->stage_ms[i];
  return GB200_OK;
}
float gb200_last_scan_kernel_ms(gb200_index *ix) { return ix ? ix->scan_kernel_ms : 0.f; }

// wait for the most recently used search context (after *_dev calls) and refresh the counters above
int gb200_sync(gb200_index *ix) {
  if (!ix) return GB200_EINVAL;
  CKI(use_device(ix));
  SearchCtx *c = nullptr;
  {
    std::lock_guard<std::mutex> g(ix->ctx_mu);
    c = ix->last_ctx;
  }
  if (!c) return GB200_OK;
  return finish_profile(ix, *c);
}
int gb200_set_profiling(gb200_index *ix, int enable) {
  if (!ix) return GB200_EINVAL;
  ix->profiling = enable != 0;
  return GB200_OK;
}
int gb200_reload_tuning(gb200_index *ix) {
  if (!ix) return GB200_EINVAL;
  std::lock_guard<std::mutex> w(ix->writer_mu);
  cudaSetDevice(ix->p.device);
  ExclusiveScope x(ix);
  ix->tune.read();
  return GB200_OK;
}

}  // extern "C"

// gb200_ivfpq_search_dev, optionally with a multi-GPU sink on its final kernel (kernels.h; called by comm.cu)
int gb_ivfpq_search_dev_sink(gb200_index *ix, int n, const float *xq_dev, int k, const gb200_search_params *sp, float *D_dev,
                             int64_t *I_dev, void *stream, const gb::PeerSink *sink, int *sink_used) {
  if (sink_used) *sink_used = 0;
  CKI(check_search_args(ix, n, xq_dev, k, sp, D_dev, I_dev));
  if (ix->kind != 0) return GB200_EINVAL;
  CKI(use_device(ix));
  SearchScope s(ix);
  if (!s.c) return GB200_ECUDA;
  // order after the caller's stream, run on the context's, and make the caller's stream wait for the result
  cudaStream_t cs = (cudaStream_t)stream;
  CKI(join_user_stream(*s.c, cs));
  s.c->sink = sink;
  s.c->sink_used = false;
  int rc = ivfpq_search_impl(ix, *s.c, n, xq_dev, true, k, sp, nullptr, 0, true, nullptr, nullptr, 0, D_dev, I_dev, true);
  s.c->sink = nullptr;
  if (sink_used) *sink_used = s.c->sink_used ? 1 : 0;
  if (rc == GB200_OK) CKI(release_to_user_stream(*s.c, cs));
  return rc;
}
