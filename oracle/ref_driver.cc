// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Thin C driver over the UNMODIFIED reference CPU engine, linked into
// oracle/_ref/liboracle_ref.so by oracle/Makefile.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// load that library.  The product (gamma_b200/csrc) never links or calls it.
//
// What it drives (reference file:line):
//   * tig_gamma::GammaIVFPQIndex::{Init,Indexing,Add,Update,Delete,Parse,Search,
//     search_preassigned}                 index/impl/gamma_index_ivfpq.cc:119-890
//   * tig_gamma::GammaFLATIndex::Search   index/impl/gamma_index_flat.cc:118-300
//   * tig_gamma::GammaSearchCondition     common/gamma_common_data.h:39-124
//     (range bitmap AND NOT deleted bitmap filter, score window, has_rank)
//   * MemoryRawVector (re-rank gather)    vector/memory_raw_vector.cc
//   * realtime::RTInvertIndex::GetIvtList realtime/realtime_invert_index.cc:77-81
// exactly the way VectorManager::Search does (vector/vector_manager.cc:433-617),
// minus the FlatBuffers/table layers that cannot be built in this container.
#include <omp.h>
#include <stdint.h>
#include <string.h>

#include <limits>
#include <map>
#include <string>
#include <vector>

#include "common/gamma_common_data.h"
#include "index/impl/gamma_index_flat.h"
#include "index/impl/gamma_index_ivfflat.h"
#include "index/impl/gamma_index_ivfpq.h"
#include "table/range_query_result.h"
#include "util/bitmap_manager.h"
#include "vector/memory_raw_vector.h"
#include "vector/raw_vector_factory.h"

#include <faiss/IndexFlat.h>
#include <faiss/VectorTransform.h>

INITIALIZE_EASYLOGGINGPP

using namespace tig_gamma;

namespace {

// GammaIndexIVFFlat::Init insists on a RocksDB raw vector unless check_vector_ is off (gamma_index_ivfflat.cc:137-151;
// the faiss-like facade index/gamma_index.cc turns it off the same way): the oracle runs it over MemoryRawVector
struct OracleIVFFlat : GammaIndexIVFFlat {
  OracleIVFFlat() { check_vector_ = false; }
};

struct OracleRef {
  int d = 0;
  bool is_ivfpq = false;
  bitmap::BitmapManager *docids_bitmap = nullptr;
  RawVector *raw = nullptr;
  GammaIVFPQIndex *ivfpq = nullptr;
  GammaFLATIndex *flat = nullptr;
  GammaIndexIVFFlat *ivfflat = nullptr;
  RetrievalModel *model = nullptr;
};

void quiet_logs() {
  static bool done = false;
  if (done) return;
  done = true;
  el::Configurations conf;
  conf.setToDefault();
  conf.setGlobally(el::ConfigurationType::Enabled, "false");
  conf.setGlobally(el::ConfigurationType::ToStandardOutput, "false");
  conf.setGlobally(el::ConfigurationType::ToFile, "false");
  el::Loggers::reconfigureAllLoggers(conf);
  el::Loggers::setDefaultConfigurations(conf, true);
}

}  // namespace

extern "C" {

void oref_set_threads(int n) { omp_set_num_threads(n); }
int oref_max_threads() { return omp_get_max_threads(); }

// retrieval_type: "IVFPQ" or "FLAT" (reflector names, index/impl/*.cc REGISTER_MODEL)
void *oref_open(const char *work_dir, int d, const char *retrieval_type,
                const char *model_json, int indexing_size, long bitmap_bits) {
  quiet_logs();
  OracleRef *o = new OracleRef;
  o->d = d;
  std::string root = std::string(work_dir);
  utils::make_dir(root.c_str());
  std::string vec_root = root + "/vectors";
  utils::make_dir(vec_root.c_str());

  VectorMetaInfo *meta = new VectorMetaInfo("gamma", d, VectorValueType::FLOAT);
  meta->with_io_ = false;
  StoreParams store_params(meta->AbsoluteName());

  o->docids_bitmap = new bitmap::BitmapManager();
  if (o->docids_bitmap->Init((uint32_t)bitmap_bits) != 0) return nullptr;

  o->raw = RawVectorFactory::Create(meta, VectorStorageType::MemoryOnly, vec_root,
                                    store_params, o->docids_bitmap);
  if (!o->raw) return nullptr;
  if (o->raw->Init("gamma", false, false) != 0) return nullptr;

  if (!strcmp(retrieval_type, "IVFPQ")) {
    o->ivfpq = new GammaIVFPQIndex();
    o->model = o->ivfpq;
    o->is_ivfpq = true;
  } else if (!strcmp(retrieval_type, "IVFFLAT")) {
    o->ivfflat = new OracleIVFFlat();
    o->model = o->ivfflat;
  } else {
    o->flat = new GammaFLATIndex();
    o->model = o->flat;
  }
  o->model->vector_ = o->raw;
  if (o->model->Init(model_json ? model_json : "", indexing_size) != 0) return nullptr;
  o->model->indexed_count_ = 0;
  return o;
}

void oref_close(void *h) {
  OracleRef *o = (OracleRef *)h;
  if (!o) return;
  if (o->ivfpq) delete o->ivfpq;
  if (o->flat) delete o->flat;
  if (o->ivfflat) delete o->ivfflat;
  if (o->raw) delete o->raw;
  if (o->docids_bitmap) delete o->docids_bitmap;
  delete o;
}

// RawVector::Add(docid, float*)  (vector/raw_vector.cc:149-156); docid == vid.
int oref_add_raw(void *h, long n, const float *x) {
  OracleRef *o = (OracleRef *)h;
  long base = o->raw->MetaInfo()->Size();
  for (long i = 0; i < n; i++) {
    if (o->raw->Add((int)(base + i), const_cast<float *>(x + i * o->d))) return -1;
  }
  return 0;
}

int oref_indexing(void *h) { return ((OracleRef *)h)->model->Indexing(); }

// the framework's AddRTVecsToIndex loop (vector/vector_manager.cc:280-382):
// feed not-yet-indexed raw vectors to model->Add in chunks.
int oref_add_to_index(void *h, long upto, int chunk) {
  OracleRef *o = (OracleRef *)h;
  long total = o->raw->MetaInfo()->Size();
  if (upto < 0 || upto > total) upto = total;
  long start = o->model->indexed_count_;
  while (start < upto) {
    long n = upto - start;
    if (n > chunk) n = chunk;
    ScopeVectors heads;
    std::vector<int> lens;
    if (o->raw->GetVectorHeader((int)start, (int)n, heads, lens)) return -1;
    for (size_t s = 0; s < heads.Size(); s++) {
      if (!o->model->Add(lens[s], heads.Get(s))) return -2;
      start += lens[s];
    }
    o->model->indexed_count_ = (int)start;
  }
  return 0;
}

// DeleteDoc: set the deleted bit, then RetrievalModel::Delete (gamma_engine.cc DelDoc path)
int oref_delete(void *h, long docid) {
  OracleRef *o = (OracleRef *)h;
  o->docids_bitmap->Set((uint32_t)docid);
  std::vector<int64_t> ids(1, docid);
  return o->model->Delete(ids);
}

int oref_update(void *h, long vid, const float *x) {
  OracleRef *o = (OracleRef *)h;
  o->raw->UpdateToStore((int)vid, (uint8_t *)x, o->d * sizeof(float));
  std::vector<int64_t> ids(1, vid);
  std::vector<const uint8_t *> vecs(1, (const uint8_t *)x);
  return o->model->Update(ids, vecs);
}

long oref_info(void *h, const char *key) {
  OracleRef *o = (OracleRef *)h;
  std::string k(key);
  if (k == "d") return o->d;
  if (k == "nraw") return (long)o->raw->MetaInfo()->Size();
  if (k == "indexed") return o->model->indexed_count_;
  if (o->ivfflat) {
    if (k == "nlist") return (long)o->ivfflat->nlist;
    if (k == "is_trained") return o->ivfflat->is_trained ? 1 : 0;
    if (k == "code_size") return (long)o->ivfflat->code_size;
    if (k == "nprobe") return (long)o->ivfflat->nprobe;
    return -1;
  }
  if (!o->ivfpq) return -1;
  GammaIVFPQIndex *ix = o->ivfpq;
  if (k == "nlist") return (long)ix->nlist;
  if (k == "M") return (long)ix->pq.M;
  if (k == "ksub") return (long)ix->pq.ksub;
  if (k == "dsub") return (long)ix->pq.dsub;
  if (k == "nbits") return (long)ix->pq.nbits;
  if (k == "code_size") return (long)ix->code_size;
  if (k == "is_trained") return ix->is_trained ? 1 : 0;
  if (k == "use_precomputed_table") return ix->use_precomputed_table;
  if (k == "by_residual") return ix->by_residual ? 1 : 0;
  if (k == "nprobe") return (long)ix->nprobe;
  if (k == "metric_ip") return ix->metric_type == faiss::METRIC_INNER_PRODUCT ? 1 : 0;
  return -1;
}

int oref_get_centroids(void *h, float *out) {
  OracleRef *o = (OracleRef *)h;
  faiss::IndexFlat *q = dynamic_cast<faiss::IndexFlat *>(o->ivfflat ? o->ivfflat->quantizer : o->ivfpq->quantizer);
  if (!q) return -1;
  memcpy(out, q->xb.data(), sizeof(float) * q->xb.size());
  return 0;
}

int oref_get_pq(void *h, float *out) {
  OracleRef *o = (OracleRef *)h;
  const std::vector<float> &c = o->ivfpq->pq.centroids;  // [M][ksub][dsub]
  memcpy(out, c.data(), sizeof(float) * c.size());
  return 0;
}

// OPQ pre-transform of the reference model (faiss::OPQMatrix : LinearTransform), if the index was created with "opq"
int oref_opq_dims(void *h, int *d_in, int *d_out, int *have_bias) {
  OracleRef *o = (OracleRef *)h;
  faiss::LinearTransform *lt = o->ivfpq ? dynamic_cast<faiss::LinearTransform *>(o->ivfpq->opq_) : nullptr;
  if (!lt) return 0;
  *d_in = lt->d_in;
  *d_out = lt->d_out;
  *have_bias = lt->have_bias ? 1 : 0;
  return 1;
}
int oref_get_opq(void *h, float *A, float *b) {
  OracleRef *o = (OracleRef *)h;
  faiss::LinearTransform *lt = o->ivfpq ? dynamic_cast<faiss::LinearTransform *>(o->ivfpq->opq_) : nullptr;
  if (!lt) return -1;
  memcpy(A, lt->A.data(), sizeof(float) * lt->A.size());
  if (lt->have_bias && b) memcpy(b, lt->b.data(), sizeof(float) * lt->b.size());
  return 0;
}

long oref_list_size(void *h, long list_no) {
  OracleRef *o = (OracleRef *)h;
  return (long)(o->ivfflat ? o->ivfflat->invlists : o->ivfpq->invlists)->list_size(list_no);
}

// raw view of one realtime inverted list: int64 ids WITH the kDelIdxMask bit
// (realtime/realtime_mem_data.h:26) and code_size bytes per posting.
int oref_get_list(void *h, long list_no, int64_t *ids, uint8_t *codes) {
  OracleRef *o = (OracleRef *)h;
  if (o->ivfflat) {  // its RTInvertIndex is private: go through the faiss InvertedLists view of it (RTInvertedLists)
    faiss::InvertedLists *il = o->ivfflat->invlists;
    size_t n = il->list_size(list_no);
    memcpy(ids, il->get_ids(list_no), n * sizeof(int64_t));
    memcpy(codes, il->get_codes(list_no), n * o->ivfflat->code_size);
    return 0;
  }
  long *ivt = nullptr;
  size_t n = 0;
  uint8_t *cds = nullptr;
  if (!o->ivfpq->rt_invert_index_ptr_->GetIvtList((size_t)list_no, ivt, n, cds)) return -1;
  memcpy(ids, ivt, n * sizeof(int64_t));
  memcpy(codes, cds, n * o->ivfpq->code_size);
  return 0;
}

// coarse quantiser exactly as GammaIVFPQIndex::Search calls it (gamma_index_ivfpq.cc:560)
int oref_coarse(void *h, int n, const float *xq, int nprobe, float *cdis, int64_t *keys) {
  OracleRef *o = (OracleRef *)h;
  (o->ivfflat ? o->ivfflat->quantizer : o->ivfpq->quantizer)->search(n, xq, nprobe, cdis, (faiss::Index::idx_t *)keys);
  return 0;
}

// Install an externally trained state, leaving the index exactly as GammaIVFPQIndex::Load does
// (gamma_index_ivfpq.cc:993-1048): quantizer holds the nlist centroids, pq.centroids set,
// is_trained, precomputed table recomputed (use_precomputed_table = 0; precompute_table()).
// Lets the bench give the reference CPU engine the SAME index the GPU path searches.
int oref_set_trained(void *h, const float *coarse, const float *pq) {
  OracleRef *o = (OracleRef *)h;
  GammaIVFPQIndex *ix = o->ivfpq;
  if (!ix) return -1;
  ix->quantizer->reset();
  ix->quantizer->add(ix->nlist, coarse);
  ix->quantizer->is_trained = true;
  memcpy(ix->pq.centroids.data(), pq, sizeof(float) * ix->pq.centroids.size());
  ix->is_trained = true;
  ix->use_precomputed_table = 0;
  if (ix->by_residual) ix->precompute_table();
  return 0;
}

// Postings straight into the realtime inverted lists: the second half of GammaIVFPQIndex::Add
// (gamma_index_ivfpq.cc:474-500) with the codes given instead of computed.  Postings must come in
// vid order so that list order equals the engine's.
int oref_inject_postings(void *h, long n, const int *list_no, const int64_t *vids, const uint8_t *codes) {
  OracleRef *o = (OracleRef *)h;
  GammaIVFPQIndex *ix = o->ivfpq;
  if (!ix) return -1;
  const size_t cs = ix->code_size;
  const long CH = 100000;
  for (long s = 0; s < n; s += CH) {
    long e = s + CH < n ? s + CH : n;
    std::map<int, std::vector<long>> new_keys;
    std::map<int, std::vector<uint8_t>> new_codes;
    for (long i = s; i < e; i++) {
      int key = list_no[i];
      new_keys[key].push_back((long)vids[i]);
      std::vector<uint8_t> &c = new_codes[key];
      c.insert(c.end(), codes + i * cs, codes + (i + 1) * cs);
    }
    if (!ix->rt_invert_index_ptr_->AddKeys(new_keys, new_codes)) return -2;
  }
  ix->indexed_vec_count_ += (int)n;
  o->model->indexed_count_ += (int)n;
  return 0;
}

static void build_filters(MultiRangeQueryResults &mr, int n_filters, const int *fmin,
                          const int *fmax, const int *fnotin, const uint8_t *const *fpass) {
  for (int f = 0; f < n_filters; f++) {
    RangeQueryResult r;
    r.SetRange(fmin[f], fmax[f]);
    r.Resize();
    int cnt = 0;
    for (int doc = fmin[f]; doc <= fmax[f]; doc++) {
      if (fpass[f][doc - fmin[f]]) {
        r.Set(doc - r.MinAligned());
        cnt++;
      }
    }
    r.SetDocNum(cnt);
    r.SetNotIn(fnotin[f] != 0);
    mr.Add(std::move(r));
  }
}

// One RetrievalModel::Search call with a hand-built GammaSearchCondition, as
// VectorManager::Search does (vector/vector_manager.cc:480-489).
//   fpass[f][doc - fmin[f]] != 0  <=>  bit set in range bitmap f.
//   keys/coarse_dis non-null => GammaIVFPQIndex::search_preassigned with the given probes.
int oref_search(void *h, int n, const float *xq, int k, const char *retrieval_json,
                int has_rank, int brute_force, float min_score, float max_score,
                int n_filters, const int *fmin, const int *fmax, const int *fnotin,
                const uint8_t *const *fpass, const int64_t *keys, const float *coarse_dis,
                int nprobe_pre, float *D, int64_t *I) {
  OracleRef *o = (OracleRef *)h;
  PerfTool perf;
  GammaSearchCondition cond(&perf);
  MultiRangeQueryResults mr;
  if (n_filters > 0) {
    build_filters(mr, n_filters, fmin, fmax, fnotin, fpass);
    cond.range_query_result = &mr;
  }
  cond.topn = k;
  cond.has_rank = has_rank != 0;
  cond.brute_force_search = brute_force != 0;
  cond.Init(min_score, max_score, o->docids_bitmap, o->raw);
  cond.retrieval_params_ = o->model->Parse(retrieval_json ? retrieval_json : "");
  // VectorResult::init pre-fills (common_query_data.h:69-82)
  for (long i = 0; i < (long)n * k; i++) {
    D[i] = 0;
    I[i] = -1;
  }
  if (keys && o->ivfpq) {
    o->ivfpq->search_preassigned(&cond, n, xq, xq, k, (const faiss::Index::idx_t *)keys,
                                 coarse_dis, D, (faiss::Index::idx_t *)I, nprobe_pre, false);
    return 0;
  }
  return o->model->Search(&cond, n, (const uint8_t *)xq, k, D, I);
}

}  // extern "C"
