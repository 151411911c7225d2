// RetrievalModel plugins for Vearch/Gamma backed by libgamma_b200.so (include/gamma_b200.h).
//
//   REGISTER_MODEL(B200IVFPQ, B200IVFPQIndex)   retrieval_type "B200IVFPQ"
//   REGISTER_MODEL(B200FLAT,  B200FLATIndex)    retrieval_type "B200FLAT"
//   REGISTER_MODEL(B200IVFFLAT, B200IVFFLATIndex) retrieval_type "B200IVFFLAT"
//
// Same pattern as the reference's own GPU model (index/impl/gpu/gamma_index_ivfpq_gpu.cc:303-305,
// 405-442): the CPU model is embedded for Init / Parse / Indexing (train) / Add (assign + PQ encode) /
// Update / Dump / Load, and ONLY Search is replaced — here by the sm_100a hot path.  Unlike the
// reference GPU model, postings are mirrored to the device incrementally (realtime), filters run
// inside the scan (pre-filter, so results equal the CPU engine's), and there is no n <= 200 limit.
//
// Compiles against the unmodified reference headers (-I<reference root>); nothing else is needed.
#pragma once
#include <pthread.h>

#include <mutex>
#include <vector>

#include "gamma_b200.h"
#include "index/impl/gamma_index_flat.h"
#include "index/impl/gamma_index_ivfflat.h"
#include "index/impl/gamma_index_ivfpq.h"

namespace tig_gamma {

// reader/writer lock that compiles as C++11 (the reference builds with -std=c++11: no std::shared_mutex)
class B200RwLock {
 public:
  B200RwLock() { pthread_rwlock_init(&l_, nullptr); }
  ~B200RwLock() { pthread_rwlock_destroy(&l_); }
  struct Shared {
    explicit Shared(B200RwLock &m) : m_(m) { pthread_rwlock_rdlock(&m_.l_); }
    ~Shared() { pthread_rwlock_unlock(&m_.l_); }
    B200RwLock &m_;
  };
  struct Exclusive {
    explicit Exclusive(B200RwLock &m) : m_(m) { pthread_rwlock_wrlock(&m_.l_); }
    ~Exclusive() { pthread_rwlock_unlock(&m_.l_); }
    B200RwLock &m_;
  };

 private:
  pthread_rwlock_t l_;
};

class B200IVFPQIndex : public GammaIVFPQIndex {
 public:
  B200IVFPQIndex();
  ~B200IVFPQIndex() override;

  int Init(const std::string &model_parameters, int indexing_size) override;
  int Indexing() override;
  bool Add(int n, const uint8_t *vec) override;
  int Update(const std::vector<int64_t> &ids, const std::vector<const uint8_t *> &vecs) override;
  int Delete(const std::vector<int64_t> &ids) override;
  int Search(RetrievalContext *retrieval_context, int n, const uint8_t *x, int k, float *distances,
             int64_t *labels) override;
  long GetTotalMemBytes() override;
  int Load(const std::string &index_dir) override;

 private:
  int PushQuantizers();
  int MirrorPostings();   // append whatever the CPU lists gained since the last call
  int MirrorRaw();        // upload raw vectors added to the store since the last call
  int MirrorList(int l);  // device copy of list l := the CPU list (after Update / CompactBucket touched it)
  int SyncDeleted();      // device live-docs bitmap := docids_bitmap_ (only changed words travel)
  int ResyncAll();        // after Load: rebuild the device index from the CPU lists (replaces dev_)
  gb200_ivfpq_params DeviceParams() const;

  // dev_mu_ (always taken before mirror_mu_): shared by every call that uses dev_, exclusive while ResyncAll replaces it —
  // the engine serves searches concurrently with Add / Update / Load.  mirror_mu_ guards the mirror bookkeeping.
  gb200_index *dev_ = nullptr;
  B200RwLock dev_mu_;
  std::mutex mirror_mu_;
  std::vector<size_t> mirrored_len_;
  long raw_mirrored_ = 0;
  long compacted_seen_ = 0;
  bool quantizers_pushed_ = false;
};

// IVFFLAT: the embedded CPU model keeps training, Add / Update and the list files; Search runs on the device over vid
// lists + the raw store.  Unlike the reference model it does not insist on a RocksDB raw vector (check_vector_ off).
class B200IVFFLATIndex : public GammaIndexIVFFlat {
 public:
  B200IVFFLATIndex();
  ~B200IVFFLATIndex() override;
  int Init(const std::string &model_parameters, int indexing_size) override;
  int Indexing() override;
  bool Add(int n, const uint8_t *vec) override;
  int Update(const std::vector<int64_t> &ids, const std::vector<const uint8_t *> &vecs) override;
  int Delete(const std::vector<int64_t> &ids);
  int Search(RetrievalContext *retrieval_context, int n, const uint8_t *x, int k, float *distances,
             int64_t *labels) override;
  long GetTotalMemBytes() override;
  int Load(const std::string &index_dir) override;

 private:
  int PushQuantizer();
  int MirrorPostings();  // append what the host lists gained
  int MirrorList(int l); // device copy of list l := the host list (RTInvertedLists view)
  int MirrorRaw();
  int SyncDeleted();
  gb200_index *dev_ = nullptr;
  B200RwLock dev_mu_;
  std::mutex mirror_mu_;
  std::vector<size_t> mirrored_len_;
  long raw_mirrored_ = 0;
  bool quantizer_pushed_ = false;
};

class B200FLATIndex : public GammaFLATIndex {
 public:
  B200FLATIndex();
  ~B200FLATIndex() override;
  int Init(const std::string &model_parameters, int indexing_size) override;
  bool Add(int n, const uint8_t *vec) override;
  int Update(const std::vector<int64_t> &ids, const std::vector<const uint8_t *> &vecs) override;
  int Delete(const std::vector<int64_t> &ids) override;
  int Search(RetrievalContext *retrieval_context, int n, const uint8_t *x, int k, float *distances,
             int64_t *labels) override;
  long GetTotalMemBytes() override;

 private:
  int MirrorRaw();
  int SyncDeleted();
  gb200_index *dev_ = nullptr;
  std::mutex mirror_mu_;
  long raw_mirrored_ = 0;
};

}  // namespace tig_gamma
