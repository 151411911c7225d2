// See b200_models.h.  Host-side glue only: every distance/scan/select runs in libgamma_b200.so.
#include "b200_models.h"

#include <string.h>

#include <algorithm>

#include <faiss/VectorTransform.h>

#include "common/gamma_common_data.h"
#include "table/range_query_result.h"
#include "vector/raw_vector.h"

namespace tig_gamma {

REGISTER_MODEL(B200IVFPQ, B200IVFPQIndex)
REGISTER_MODEL(B200FLAT, B200FLATIndex)
REGISTER_MODEL(B200IVFFLAT, B200IVFFLATIndex)

namespace {

int DeviceOrdinal() {
  const char *e = getenv("GB200_DEVICE");
  return e ? atoi(e) : 0;
}

// range filters of the request -> C-ABI descriptors (RangeQueryResult, table/range_query_result.h:24-160)
void CollectFilters(GammaSearchCondition *cond, std::vector<gb200_range_filter> *out) {
  MultiRangeQueryResults *mr = cond ? cond->range_query_result : nullptr;
  if (!mr) return;
  const RangeQueryResult *all = mr->GetAllResult();
  for (size_t i = 0; i < mr->Size(); i++) {
    RangeQueryResult &r = const_cast<RangeQueryResult &>(all[i]);
    gb200_range_filter f;
    f.min_doc = r.Min();
    f.max_doc = r.Max();
    f.min_aligned = r.MinAligned();
    f.not_in = r.NotIn() ? 1 : 0;
    f.bitmap = reinterpret_cast<const uint8_t *>(r.Ref());
    out->push_back(f);
  }
}

// NB: a request with range_query_result != nullptr but zero results matches nothing
// (MultiRangeQueryResults::Has returns false when all_results_ is empty).
bool MatchesNothing(GammaSearchCondition *cond) {
  return cond && cond->range_query_result && cond->range_query_result->Size() == 0;
}

void FillEmpty(int n, int k, bool ip, float *distances, int64_t *labels) {
  for (long i = 0; i < (long)n * k; i++) {
    distances[i] = ip ? -std::numeric_limits<float>::max() : std::numeric_limits<float>::max();
    labels[i] = -1;
  }
}

int UploadRawRange(gb200_index *dev, RawVector *raw, long from, long to) {
  while (from < to) {
    ScopeVectors heads;
    std::vector<int> lens;
    long n = std::min<long>(to - from, 1 << 20);
    if (raw->GetVectorHeader((int)from, (int)n, heads, lens)) return -1;
    for (size_t s = 0; s < heads.Size(); s++) {
      int rc = gb200_upload_raw(dev, from, lens[s], reinterpret_cast<const float *>(heads.Get(s)));
      if (rc) return rc;
      from += lens[s];
    }
  }
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------ IVFPQ
B200IVFPQIndex::B200IVFPQIndex() {}
B200IVFPQIndex::~B200IVFPQIndex() {
  if (dev_) gb200_destroy(dev_);
}

// the validity bitmap is indexed by doc id while postings carry vids: the two coincide only for single-vector fields
// (VIDMgr::VID2DocID is the identity unless multi_vids_, vector/raw_vector_common.h:44-147)
static bool MultiVidStore(VectorReader *v) {
  RawVector *raw = dynamic_cast<RawVector *>(v);
  return raw && raw->VidMgr() && raw->VidMgr()->MultiVids();
}

gb200_ivfpq_params B200IVFPQIndex::DeviceParams() const {
  gb200_ivfpq_params p;
  memset(&p, 0, sizeof(p));
  p.device = DeviceOrdinal();
  p.d = this->d;
  p.raw_d = vector_->MetaInfo()->Dimension();
  p.nlist = (int)this->nlist;
  p.nsubvector = (int)this->pq.M;
  p.nbits = (int)this->pq.nbits;
  p.metric = metric_type_ == DistanceComputeType::INNER_PRODUCT ? GB200_METRIC_INNER_PRODUCT : GB200_METRIC_L2;
  p.nprobe = (int)this->nprobe;
  p.store_raw = 1;
  return p;
}

int B200IVFPQIndex::Init(const std::string &model_parameters, int indexing_size) {
  int ret = GammaIVFPQIndex::Init(model_parameters, indexing_size);
  if (ret) return ret;
  if (quantizer_type_ != 0) {
    LOG(ERROR) << "B200IVFPQ: the hnsw coarse quantizer is not supported";
    return -1;
  }
  if (MultiVidStore(vector_)) {
    LOG(ERROR) << "B200IVFPQ: multi-vid vector fields are not supported (filters are applied by vid)";
    return -1;
  }
  gb200_ivfpq_params p = DeviceParams();
  int rc = gb200_ivfpq_create(&p, &dev_);
  if (rc) {
    LOG(ERROR) << "gb200_ivfpq_create failed: " << rc << " " << gb200_last_error();
    return -1;
  }
  mirrored_len_.assign(this->nlist, 0);
  return 0;
}

int B200IVFPQIndex::PushQuantizers() {
  faiss::IndexFlat *flat = dynamic_cast<faiss::IndexFlat *>(this->quantizer);
  if (!flat || !this->is_trained) return -1;
  int rc = gb200_ivfpq_set_quantizers(dev_, flat->xb.data(), this->pq.centroids.data());
  if (rc == 0 && opq_ != nullptr) {  // model parameter "opq": the trained OPQMatrix travels with the quantizers
    faiss::LinearTransform *lt = dynamic_cast<faiss::LinearTransform *>(opq_);
    if (!lt) return -1;
    rc = gb200_ivfpq_set_opq(dev_, lt->d_in, lt->d_out, lt->A.data(), lt->have_bias ? lt->b.data() : nullptr);
  }
  if (rc == 0) quantizers_pushed_ = true;
  return rc;
}

int B200IVFPQIndex::Indexing() {
  int ret = GammaIVFPQIndex::Indexing();  // faiss::IndexIVFPQ::train on the host (gamma_index_ivfpq.cc:272-354)
  if (ret) return ret;
  B200RwLock::Shared dl(dev_mu_);
  std::lock_guard<std::mutex> g(mirror_mu_);
  return PushQuantizers() ? -1 : 0;
}

int B200IVFPQIndex::MirrorList(int l) {
  long *ids = nullptr;
  size_t len = 0;
  uint8_t *cds = nullptr;
  if (!rt_invert_index_ptr_->GetIvtList(l, ids, len, cds)) len = 0;
  // idx_array_ words as they are (bit 63 = kDelIdxMask): gb200_ivfpq_replace_list takes the reference layout
  int rc = gb200_ivfpq_replace_list(dev_, l, (int64_t)len, reinterpret_cast<const int64_t *>(ids), cds);
  if (rc == 0) mirrored_len_[l] = len;
  return rc;
}

int B200IVFPQIndex::SyncDeleted() {
  RawVector *raw = dynamic_cast<RawVector *>(vector_);
  bitmap::BitmapManager *bm = raw ? raw->Bitmap() : nullptr;
  if (!bm) return 0;
  return gb200_upload_deleted_bitmap(dev_, reinterpret_cast<const uint8_t *>(bm->Bitmap()), bm->BitSize());
}

int B200IVFPQIndex::MirrorPostings() {
  std::vector<int32_t> list_no;
  std::vector<int64_t> vids;
  std::vector<uint8_t> codes;
  for (size_t l = 0; l < this->nlist; l++) {
    long *ids = nullptr;
    size_t len = 0;
    uint8_t *cds = nullptr;
    if (!rt_invert_index_ptr_->GetIvtList(l, ids, len, cds)) continue;
    for (size_t j = mirrored_len_[l]; j < len; j++) {
      list_no.push_back((int32_t)l);
      vids.push_back(ids[j] & realtime::kRecoverIdxMask);
      codes.insert(codes.end(), cds + j * code_size, cds + (j + 1) * code_size);
    }
    mirrored_len_[l] = len;
  }
  if (list_no.empty()) return 0;
  return gb200_ivfpq_append(dev_, (int64_t)list_no.size(), list_no.data(), vids.data(), codes.data());
}

int B200IVFPQIndex::MirrorRaw() {
  RawVector *raw = dynamic_cast<RawVector *>(vector_);
  long total = (long)raw->MetaInfo()->Size();
  if (total <= raw_mirrored_) return 0;
  int rc = UploadRawRange(dev_, raw, raw_mirrored_, total);
  if (rc == 0) raw_mirrored_ = total;
  return rc;
}

int B200IVFPQIndex::ResyncAll() {
  // rebuild from scratch after Load.  Caller holds dev_mu_ EXCLUSIVELY (no search is inside the library) and mirror_mu_.
  gb200_ivfpq_params p = DeviceParams();
  if (dev_) gb200_destroy(dev_);
  dev_ = nullptr;
  if (gb200_ivfpq_create(&p, &dev_)) return -1;
  mirrored_len_.assign(this->nlist, 0);
  raw_mirrored_ = 0;
  if (PushQuantizers()) return -1;
  // one bulk append of the alive postings (fast path for a big index), then the lists that hold dead slots
  // (kDelIdxMask: the slot stays so that positions match the host's) are rewritten with the flags in place
  std::vector<int32_t> list_no;
  std::vector<int64_t> vids;
  std::vector<uint8_t> codes;
  std::vector<int> with_dead;
  for (size_t l = 0; l < this->nlist; l++) {
    long *ids = nullptr;
    size_t len = 0;
    uint8_t *cds = nullptr;
    if (!rt_invert_index_ptr_->GetIvtList(l, ids, len, cds)) continue;
    bool dead = false;
    for (size_t j = 0; j < len; j++) dead |= (ids[j] & realtime::kDelIdxMask) != 0;
    if (dead) {
      with_dead.push_back((int)l);
      continue;
    }
    for (size_t j = 0; j < len; j++) {
      list_no.push_back((int32_t)l);
      vids.push_back(ids[j]);
      codes.insert(codes.end(), cds + j * code_size, cds + (j + 1) * code_size);
    }
    mirrored_len_[l] = len;
  }
  if (!list_no.empty() && gb200_ivfpq_append(dev_, (int64_t)list_no.size(), list_no.data(), vids.data(), codes.data()))
    return -1;
  for (int l : with_dead)
    if (MirrorList(l)) return -1;
  if (SyncDeleted()) return -1;
  compacted_seen_ = rt_invert_index_ptr_->cur_ptr_->cur_invert_ptr_->compacted_num_;
  return MirrorRaw();
}

bool B200IVFPQIndex::Add(int n, const uint8_t *vec) {
  if (!GammaIVFPQIndex::Add(n, vec)) return false;  // assign + residual + pq.compute_codes + AddKeys on the host
  B200RwLock::Shared dl(dev_mu_);
  std::lock_guard<std::mutex> g(mirror_mu_);
  if (!quantizers_pushed_ && PushQuantizers()) return false;
  return MirrorPostings() == 0;
}

int B200IVFPQIndex::Update(const std::vector<int64_t> &ids, const std::vector<const uint8_t *> &vecs) {
  // lists the update may touch, taken BEFORE the host applies it: where every vid lives now
  std::vector<int> touched;
  {
    realtime::RTInvertBucketData *cur = rt_invert_index_ptr_->cur_ptr_->cur_invert_ptr_;
    for (size_t i = 0; i < ids.size(); i++) {
      if (ids[i] < 0 || (size_t)ids[i] >= cur->nids_) continue;
      long loc = cur->vid_bucket_no_pos_[ids[i]];
      if (loc != -1) touched.push_back((int)(loc >> 32));
    }
  }
  int ret = GammaIVFPQIndex::Update(ids, vecs);  // re-assign, re-encode, RealTimeMemData::Update, CompactIfNeed
  if (ret) return ret;
  B200RwLock::Shared dl(dev_mu_);
  std::lock_guard<std::mutex> g(mirror_mu_);
  realtime::RTInvertBucketData *cur = rt_invert_index_ptr_->cur_ptr_->cur_invert_ptr_;
  RawVector *raw = dynamic_cast<RawVector *>(vector_);
  for (size_t i = 0; i < ids.size(); i++) {  // ... and where it lives afterwards
    long vid = ids[i];
    if (vid < 0 || (size_t)vid >= cur->nids_) continue;
    long loc = cur->vid_bucket_no_pos_[vid];
    if (loc != -1) touched.push_back((int)(loc >> 32));
    // the raw vector changed too (re-rank reads it)
    ScopeVector sv;
    raw->GetVector(vid, sv);
    if (sv.Get() && gb200_upload_raw(dev_, vid, 1, reinterpret_cast<const float *>(sv.Get()))) return -1;
  }
  if (cur->compacted_num_ != compacted_seen_) {  // CompactBucket rewrote some lists: the ones that got shorter
    if (SyncDeleted()) return -1;
    for (size_t l = 0; l < this->nlist; l++)
      if ((size_t)cur->retrieve_idx_pos_[l] < mirrored_len_[l]) touched.push_back((int)l);
    compacted_seen_ = cur->compacted_num_;
  }
  std::sort(touched.begin(), touched.end());
  touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
  // every touched list is replaced by the host's copy (dead flags included): searches keep running, the swap is a
  // publication of the new extent
  for (int l : touched)
    if (MirrorList(l)) return -1;
  return 0;
}

int B200IVFPQIndex::Delete(const std::vector<int64_t> &ids) {
  GammaIVFPQIndex::Delete(ids);
  // liveness is the deleted-docs bitmap (set by the engine before this call); mirror the bits
  B200RwLock::Shared dl(dev_mu_);
  return gb200_set_deleted(dev_, ids.data(), (int64_t)ids.size(), 1) ? -1 : 0;
}

int B200IVFPQIndex::Load(const std::string &index_dir) {
  int ret = GammaIVFPQIndex::Load(index_dir);
  if (ret < 0) return ret;
  B200RwLock::Exclusive dl(dev_mu_);  // the device index is replaced: no search may be inside it
  std::lock_guard<std::mutex> g(mirror_mu_);
  if (this->is_trained && ResyncAll()) return -1;
  return ret;
}

long B200IVFPQIndex::GetTotalMemBytes() { return GammaIVFPQIndex::GetTotalMemBytes() + (dev_ ? gb200_mem_bytes(dev_) : 0); }

int B200IVFPQIndex::Search(RetrievalContext *retrieval_context, int n, const uint8_t *x, int k, float *distances,
                           int64_t *labels) {
  IVFPQRetrievalParameters *rp = dynamic_cast<IVFPQRetrievalParameters *>(retrieval_context->RetrievalParams());
  IVFPQRetrievalParameters dflt;
  if (rp == nullptr) rp = &dflt;
  GammaSearchCondition *cond = dynamic_cast<GammaSearchCondition *>(retrieval_context);
  gb200_search_params sp;
  sp.metric = rp->GetDistanceComputeType() == DistanceComputeType::INNER_PRODUCT ? GB200_METRIC_INNER_PRODUCT
                                                                                 : GB200_METRIC_L2;
  sp.nprobe = rp->Nprobe();
  sp.recall_num = rp->RecallNum();
  sp.has_rank = cond ? (cond->has_rank ? 1 : 0) : 1;
  sp.min_score = cond ? cond->min_score : -std::numeric_limits<float>::max();
  sp.max_score = cond ? cond->max_score : std::numeric_limits<float>::max();
  if (k <= 0) return 0;  // reference logs and returns (gamma_index_ivfpq.cc:753-756)
  if (MatchesNothing(cond)) {
    FillEmpty(n, k, sp.metric == GB200_METRIC_INNER_PRODUCT, distances, labels);
    return 0;
  }
  B200RwLock::Shared dl(dev_mu_);  // held for the whole device search
  {
    std::lock_guard<std::mutex> g(mirror_mu_);
    if (MirrorRaw()) return -1;  // the store grows ahead of the index (AddToStore, gamma_engine.cc:651)
    // the reference tests docids_bitmap_ live, and some engine paths set bits without calling Delete()
    // (DelDocByQuery, search/gamma_engine.cc:866): bring the device bitmap in line (a memcmp when nothing changed)
    if (SyncDeleted()) return -1;
  }
  std::vector<gb200_range_filter> filters;
  CollectFilters(cond, &filters);
  const float *xq = reinterpret_cast<const float *>(x);
  std::vector<float> padded;
  int raw_d = vector_->MetaInfo()->Dimension();
  if (this->d > raw_d) {  // support_indivisible_nsubvector: zero-pad the queries (ConvertVectorDim)
    padded.assign((size_t)n * this->d, 0.f);
    for (int i = 0; i < n; i++) memcpy(&padded[(size_t)i * this->d], xq + (size_t)i * raw_d, raw_d * sizeof(float));
    xq = padded.data();
  }
  int rc;
  if ((cond && cond->brute_force_search) || !this->is_trained) {
    rc = gb200_flat_search(dev_, n, reinterpret_cast<const float *>(x), k, &sp, filters.data(), (int)filters.size(),
                           distances, labels);
  } else {
    rc = gb200_ivfpq_search(dev_, n, xq, k, &sp, filters.data(), (int)filters.size(), distances, labels);
  }
  if (rc) LOG(ERROR) << "gb200 search failed: " << rc << " " << gb200_last_error();
  return rc;
}

// ------------------------------------------------------------------------------------ IVFFLAT
B200IVFFLATIndex::B200IVFFLATIndex() { check_vector_ = false; }  // any RawVector will do: the device reads its own copy
B200IVFFLATIndex::~B200IVFFLATIndex() {
  if (dev_) gb200_destroy(dev_);
}

int B200IVFFLATIndex::Init(const std::string &model_parameters, int indexing_size) {
  int ret = GammaIndexIVFFlat::Init(model_parameters, indexing_size);
  if (ret) return ret;
  if (MultiVidStore(vector_)) {
    LOG(ERROR) << "B200IVFFLAT: multi-vid vector fields are not supported (filters are applied by vid)";
    return -1;
  }
  int metric = this->metric_type == faiss::METRIC_INNER_PRODUCT ? GB200_METRIC_INNER_PRODUCT : GB200_METRIC_L2;
  if (gb200_ivfflat_create(DeviceOrdinal(), (int)this->d, (int)this->nlist, metric, (int)this->nprobe, &dev_)) {
    LOG(ERROR) << "gb200_ivfflat_create failed: " << gb200_last_error();
    return -1;
  }
  mirrored_len_.assign(this->nlist, 0);
  return 0;
}

int B200IVFFLATIndex::PushQuantizer() {
  faiss::IndexFlat *flat = dynamic_cast<faiss::IndexFlat *>(this->quantizer);
  if (!flat || !this->is_trained) return -1;
  int rc = gb200_ivfflat_set_quantizer(dev_, flat->xb.data());
  if (rc == 0) quantizer_pushed_ = true;
  return rc;
}

int B200IVFFLATIndex::Indexing() {
  int ret = GammaIndexIVFFlat::Indexing();
  if (ret) return ret;
  B200RwLock::Shared dl(dev_mu_);
  std::lock_guard<std::mutex> g(mirror_mu_);
  return PushQuantizer() ? -1 : 0;
}

int B200IVFFLATIndex::MirrorRaw() {
  RawVector *raw = dynamic_cast<RawVector *>(vector_);
  long total = (long)raw->MetaInfo()->Size();
  if (total <= raw_mirrored_) return 0;
  int rc = UploadRawRange(dev_, raw, raw_mirrored_, total);
  if (rc == 0) raw_mirrored_ = total;
  return rc;
}

int B200IVFFLATIndex::SyncDeleted() {
  RawVector *raw = dynamic_cast<RawVector *>(vector_);
  bitmap::BitmapManager *bm = raw ? raw->Bitmap() : nullptr;
  if (!bm) return 0;
  return gb200_upload_deleted_bitmap(dev_, reinterpret_cast<const uint8_t *>(bm->Bitmap()), bm->BitSize());
}

// the model's RTInvertIndex is private: the lists are read through the faiss InvertedLists view of it (RTInvertedLists)
int B200IVFFLATIndex::MirrorPostings() {
  std::vector<int32_t> list_no;
  std::vector<int64_t> vids;
  for (size_t l = 0; l < this->nlist; l++) {
    size_t len = this->invlists->list_size(l);
    if (len <= mirrored_len_[l]) continue;
    const faiss::Index::idx_t *ids = this->invlists->get_ids(l);
    for (size_t j = mirrored_len_[l]; j < len; j++) {
      list_no.push_back((int32_t)l);
      vids.push_back(ids[j] & realtime::kRecoverIdxMask);
    }
    mirrored_len_[l] = len;
  }
  if (list_no.empty()) return 0;
  return gb200_ivfflat_append(dev_, (int64_t)list_no.size(), list_no.data(), vids.data());
}

int B200IVFFLATIndex::MirrorList(int l) {
  size_t len = this->invlists->list_size(l);
  const faiss::Index::idx_t *ids = this->invlists->get_ids(l);
  std::vector<uint8_t> codes(len * 4, 0);  // the device lists carry no codes: 4 dummy bytes per posting
  int rc = gb200_ivfpq_replace_list(dev_, l, (int64_t)len, reinterpret_cast<const int64_t *>(ids), codes.data());
  if (rc == 0) mirrored_len_[l] = len;
  return rc;
}

bool B200IVFFLATIndex::Add(int n, const uint8_t *vec) {
  if (!GammaIndexIVFFlat::Add(n, vec)) return false;  // quantizer->assign + AddKeys on the host
  B200RwLock::Shared dl(dev_mu_);
  std::lock_guard<std::mutex> g(mirror_mu_);
  if (!quantizer_pushed_ && PushQuantizer()) return false;
  if (MirrorRaw()) return false;  // the scan reads the vectors from the device raw store
  return MirrorPostings() == 0;
}

int B200IVFFLATIndex::Update(const std::vector<int64_t> &ids, const std::vector<const uint8_t *> &vecs) {
  int ret = GammaIndexIVFFlat::Update(ids, vecs);  // re-assign, RealTimeMemData::Update, CompactIfNeed
  if (ret) return ret;
  B200RwLock::Shared dl(dev_mu_);
  std::lock_guard<std::mutex> g(mirror_mu_);
  RawVector *raw = dynamic_cast<RawVector *>(vector_);
  for (size_t i = 0; i < ids.size(); i++) {
    ScopeVector sv;
    raw->GetVector(ids[i], sv);
    if (sv.Get() && gb200_upload_raw(dev_, ids[i], 1, reinterpret_cast<const float *>(sv.Get()))) return -1;
  }
  // RealTimeMemData::Update on the device: the old posting keeps its slot with the kDelIdxMask flag, the new one is
  // appended to the list the quantizer assigns (same list: nothing moves — the vector itself lives in the raw store)
  const uint8_t dummy[4] = {0, 0, 0, 0};
  for (size_t i = 0; i < ids.size(); i++) {
    faiss::Index::idx_t idx = -1;
    quantizer->assign(1, reinterpret_cast<const float *>(vecs[i]), &idx);
    if (idx < 0) continue;
    if (gb200_ivfpq_update(dev_, ids[i], (int32_t)idx, dummy)) return -1;
  }
  // CompactIfNeed may have rewritten lists on the host: those whose length now differs from the device's are replaced
  if (SyncDeleted()) return -1;
  std::vector<int64_t> dev_len(this->nlist);
  if (gb200_ivfpq_list_sizes(dev_, dev_len.data())) return -1;
  for (size_t l = 0; l < this->nlist; l++) {
    if ((int64_t)this->invlists->list_size(l) != dev_len[l]) {
      if (MirrorList((int)l)) return -1;
    } else {
      mirrored_len_[l] = (size_t)dev_len[l];
    }
  }
  return 0;
}

int B200IVFFLATIndex::Delete(const std::vector<int64_t> &ids) {
  GammaIndexIVFFlat::Delete(ids);
  B200RwLock::Shared dl(dev_mu_);
  return gb200_set_deleted(dev_, ids.data(), (int64_t)ids.size(), 1) ? -1 : 0;
}

int B200IVFFLATIndex::Load(const std::string &index_dir) {
  int ret = GammaIndexIVFFlat::Load(index_dir);
  if (ret < 0) return ret;
  B200RwLock::Exclusive dl(dev_mu_);
  std::lock_guard<std::mutex> g(mirror_mu_);
  if (!this->is_trained) return ret;
  if (PushQuantizer() || MirrorRaw() || SyncDeleted()) return -1;
  for (size_t l = 0; l < this->nlist; l++)
    if (MirrorList((int)l)) return -1;
  return ret;
}

long B200IVFFLATIndex::GetTotalMemBytes() { return dev_ ? gb200_mem_bytes(dev_) : 0; }

int B200IVFFLATIndex::Search(RetrievalContext *retrieval_context, int n, const uint8_t *x, int k, float *distances,
                             int64_t *labels) {
  IVFFlatRetrievalParameters *rp = dynamic_cast<IVFFlatRetrievalParameters *>(retrieval_context->RetrievalParams());
  GammaSearchCondition *cond = dynamic_cast<GammaSearchCondition *>(retrieval_context);
  gb200_search_params sp;
  memset(&sp, 0, sizeof(sp));
  DistanceComputeType t = rp ? rp->GetDistanceComputeType()
                             : (this->metric_type == faiss::METRIC_INNER_PRODUCT ? DistanceComputeType::INNER_PRODUCT
                                                                                : DistanceComputeType::L2);
  sp.metric = t == DistanceComputeType::INNER_PRODUCT ? GB200_METRIC_INNER_PRODUCT : GB200_METRIC_L2;
  sp.nprobe = rp ? rp->Nprobe() : -1;  // <= 0: the model's nprobe
  sp.min_score = cond ? cond->min_score : -std::numeric_limits<float>::max();
  sp.max_score = cond ? cond->max_score : std::numeric_limits<float>::max();
  if (k <= 0) return 0;
  if (MatchesNothing(cond)) {
    FillEmpty(n, k, sp.metric == GB200_METRIC_INNER_PRODUCT, distances, labels);
    return 0;
  }
  B200RwLock::Shared dl(dev_mu_);
  {
    std::lock_guard<std::mutex> g(mirror_mu_);
    if (MirrorRaw()) return -1;
    if (SyncDeleted()) return -1;
  }
  std::vector<gb200_range_filter> filters;
  CollectFilters(cond, &filters);
  int rc = gb200_ivfflat_search(dev_, n, reinterpret_cast<const float *>(x), k, &sp, filters.data(), (int)filters.size(),
                                distances, labels);
  if (rc) LOG(ERROR) << "gb200 ivfflat search failed: " << rc << " " << gb200_last_error();
  return rc;
}

// ------------------------------------------------------------------------------------ FLAT
B200FLATIndex::B200FLATIndex() {}
B200FLATIndex::~B200FLATIndex() {
  if (dev_) gb200_destroy(dev_);
}

int B200FLATIndex::Init(const std::string &model_parameters, int indexing_size) {
  int ret = GammaFLATIndex::Init(model_parameters, indexing_size);
  if (ret) return ret;
  if (MultiVidStore(vector_)) {
    LOG(ERROR) << "B200FLAT: multi-vid vector fields are not supported (filters are applied by vid)";
    return -1;
  }
  int metric = metric_type_ == DistanceComputeType::INNER_PRODUCT ? GB200_METRIC_INNER_PRODUCT : GB200_METRIC_L2;
  if (gb200_flat_create(DeviceOrdinal(), vector_->MetaInfo()->Dimension(), metric, &dev_)) {
    LOG(ERROR) << "gb200_flat_create failed: " << gb200_last_error();
    return -1;
  }
  return 0;
}

int B200FLATIndex::SyncDeleted() {
  RawVector *raw = dynamic_cast<RawVector *>(vector_);
  bitmap::BitmapManager *bm = raw ? raw->Bitmap() : nullptr;
  if (!bm) return 0;
  return gb200_upload_deleted_bitmap(dev_, reinterpret_cast<const uint8_t *>(bm->Bitmap()), bm->BitSize());
}

int B200FLATIndex::MirrorRaw() {
  RawVector *raw = dynamic_cast<RawVector *>(vector_);
  long total = (long)raw->MetaInfo()->Size();
  if (total <= raw_mirrored_) return 0;
  int rc = UploadRawRange(dev_, raw, raw_mirrored_, total);
  if (rc == 0) raw_mirrored_ = total;
  return rc;
}

bool B200FLATIndex::Add(int n, const uint8_t *vec) {
  std::lock_guard<std::mutex> g(mirror_mu_);
  return MirrorRaw() == 0;
}

int B200FLATIndex::Update(const std::vector<int64_t> &ids, const std::vector<const uint8_t *> &vecs) {
  for (size_t i = 0; i < ids.size(); i++)
    if (gb200_upload_raw(dev_, ids[i], 1, reinterpret_cast<const float *>(vecs[i]))) return -1;
  return 0;
}

int B200FLATIndex::Delete(const std::vector<int64_t> &ids) {
  return gb200_set_deleted(dev_, ids.data(), (int64_t)ids.size(), 1) ? -1 : 0;
}

long B200FLATIndex::GetTotalMemBytes() { return dev_ ? gb200_mem_bytes(dev_) : 0; }

int B200FLATIndex::Search(RetrievalContext *retrieval_context, int n, const uint8_t *x, int k, float *distances,
                          int64_t *labels) {
  FlatRetrievalParameters *rp = dynamic_cast<FlatRetrievalParameters *>(retrieval_context->RetrievalParams());
  GammaSearchCondition *cond = dynamic_cast<GammaSearchCondition *>(retrieval_context);
  gb200_search_params sp;
  memset(&sp, 0, sizeof(sp));
  DistanceComputeType t = rp ? rp->GetDistanceComputeType() : DistanceComputeType::L2;
  sp.metric = t == DistanceComputeType::INNER_PRODUCT ? GB200_METRIC_INNER_PRODUCT : GB200_METRIC_L2;
  sp.min_score = cond ? cond->min_score : -std::numeric_limits<float>::max();
  sp.max_score = cond ? cond->max_score : std::numeric_limits<float>::max();
  if (k <= 0) return 0;
  if (MatchesNothing(cond)) {
    FillEmpty(n, k, sp.metric == GB200_METRIC_INNER_PRODUCT, distances, labels);
    return 0;
  }
  {
    std::lock_guard<std::mutex> g(mirror_mu_);
    if (MirrorRaw()) return -1;
    if (SyncDeleted()) return -1;
  }
  std::vector<gb200_range_filter> filters;
  CollectFilters(cond, &filters);
  int rc = gb200_flat_search(dev_, n, reinterpret_cast<const float *>(x), k, &sp, filters.data(), (int)filters.size(),
                             distances, labels);
  if (rc) LOG(ERROR) << "gb200 flat search failed: " << rc << " " << gb200_last_error();
  return rc;
}

}  // namespace tig_gamma
