// RetrievalModel plugins for Vearch/Gamma backed by libgamma_b200.so (include/gamma_b200.h).
//
//   REGISTER_MODEL(B200IVFPQ, B200IVFPQIndex)   retrieval_type "B200IVFPQ"
//   REGISTER_MODEL(B200FLAT,  B200FLATIndex)    retrieval_type "B200FLAT"
//
// Same pattern as the reference's own GPU model (index/impl/gpu/gamma_index_ivfpq_gpu.cc:303-305,
// 405-442): the CPU model is embedded for Init / Parse / Indexing (train) / Add (assign + PQ encode) /
// Update / Dump / Load, and ONLY Search is replaced — here by the sm_100a hot path.  Unlike the
// reference GPU model, postings are mirrored to the device incrementally (realtime), filters run
// inside the scan (pre-filter, so results equal the CPU engine's), and there is no n <= 200 limit.
//
// Compiles against the unmodified reference headers (-I<reference root>); nothing else is needed.
#pragma once
#include <mutex>
#include <vector>

#include "gamma_b200.h"
#include "index/impl/gamma_index_flat.h"
#include "index/impl/gamma_index_ivfpq.h"

namespace tig_gamma {

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
  int ResyncAll();        // after Load / compaction: rebuild the device lists from the CPU lists

  gb200_index *dev_ = nullptr;
  std::mutex mirror_mu_;
  std::vector<size_t> mirrored_len_;
  long raw_mirrored_ = 0;
  long compacted_seen_ = 0;
  bool quantizers_pushed_ = false;
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
  gb200_index *dev_ = nullptr;
  std::mutex mirror_mu_;
  long raw_mirrored_ = 0;
};

}  // namespace tig_gamma
