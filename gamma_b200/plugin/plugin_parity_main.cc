// Drop-in proof: the reference's own model ("IVFPQ"/"FLAT") and the B200 plugin ("B200IVFPQ"/"B200FLAT")
// are obtained from the SAME reflector (index/reflector.h:50-57, exactly what VectorManager does,
// vector/vector_manager.cc:161-195), fed the same RawVector / deleted bitmap / Add / Update / Delete
// stream, and searched with the same GammaSearchCondition; outputs are compared.
// Built by `make -C oracle plugin` against the unmodified reference sources; runs on the GPU box.
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <random>
#include <string>
#include <vector>

#include "b200_models.h"
#include "common/gamma_common_data.h"
#include "index/reflector.h"
#include "table/range_query_result.h"
#include "util/bitmap_manager.h"
#include "vector/raw_vector_factory.h"

#include <faiss/IndexFlat.h>

INITIALIZE_EASYLOGGINGPP
using namespace tig_gamma;

static void quiet() {
  el::Configurations conf;
  conf.setToDefault();
  conf.setGlobally(el::ConfigurationType::Enabled, "false");
  el::Loggers::reconfigureAllLoggers(conf);
  el::Loggers::setDefaultConfigurations(conf, true);
}

struct Cmp {
  long slots = 0, same_id = 0, filled_mismatch = 0;
  double max_rel = 0;
};

static Cmp compare(int n, int k, const float *D0, const int64_t *I0, const float *D1, const int64_t *I1) {
  Cmp c;
  for (long i = 0; i < (long)n * k; i++) {
    c.slots++;
    if (I0[i] == I1[i]) {
      c.same_id++;
      if (I0[i] >= 0) {
        double rel = fabs((double)D0[i] - D1[i]) / fmax(fabs((double)D0[i]), 1e-6);
        if (rel > c.max_rel) c.max_rel = rel;
      }
    }
    if ((I0[i] < 0) != (I1[i] < 0)) c.filled_mismatch++;
  }
  return c;
}

int main(int argc, char **argv) {
  quiet();
  const int N = argc > 1 ? atoi(argv[1]) : 60000, d = 128, nlist = 256, M = 32, nq = 128, k = 10;
  omp_set_num_threads(16);
  std::mt19937 rng(123);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> centres(256 * d), xb((size_t)N * d), xq((size_t)nq * d);
  for (auto &v : centres) v = nd(rng);
  auto sample = [&](float *out) {
    int c = rng() % 256;
    for (int j = 0; j < d; j++) out[j] = centres[c * d + j] + 0.3f * nd(rng);
  };
  for (int i = 0; i < N; i++) sample(&xb[(size_t)i * d]);
  for (int i = 0; i < nq; i++) sample(&xq[(size_t)i * d]);

  utils::make_dir("/tmp/b200_plugin_parity");
  utils::make_dir("/tmp/b200_plugin_parity/vectors");
  VectorMetaInfo *meta = new VectorMetaInfo("gamma", d, VectorValueType::FLOAT);
  meta->with_io_ = false;
  StoreParams sp(meta->AbsoluteName());
  bitmap::BitmapManager *bm = new bitmap::BitmapManager();
  bm->Init(4 * N);
  RawVector *raw = RawVectorFactory::Create(meta, VectorStorageType::MemoryOnly, "/tmp/b200_plugin_parity/vectors", sp, bm);
  if (!raw || raw->Init("gamma", false, false)) {
    printf("{\"error\":\"raw vector init failed\"}\n");
    return 2;
  }

  const char *json = "{\"ncentroids\":256,\"nsubvector\":32,\"metric_type\":\"L2\",\"nprobe\":16}";
  RetrievalModel *cpu = reflector().GetNewModel("IVFPQ");
  RetrievalModel *gpu = reflector().GetNewModel("B200IVFPQ");
  RetrievalModel *cpu_flat = reflector().GetNewModel("FLAT");
  RetrievalModel *gpu_flat = reflector().GetNewModel("B200FLAT");
  if (!cpu || !gpu || !cpu_flat || !gpu_flat) {
    printf("{\"error\":\"reflector did not return the models\"}\n");
    return 3;
  }
  for (RetrievalModel *m : {cpu, gpu, cpu_flat, gpu_flat}) m->vector_ = raw;
  const int first = N - 5000;  // docs present when the index is trained (= indexing_size)
  if (cpu->Init(json, first)) {
    printf("{\"error\":\"cpu model Init failed\"}\n");
    return 4;
  }
  if (gpu->Init(json, first)) {
    printf("{\"error\":\"B200IVFPQ Init failed: %s\"}\n", gb200_last_error());
    return 4;
  }
  if (cpu_flat->Init("{\"metric_type\":\"L2\"}", N) || gpu_flat->Init("{\"metric_type\":\"L2\"}", N)) {
    printf("{\"error\":\"flat Init failed: %s\"}\n", gb200_last_error());
    return 4;
  }

  // the engine adds to the store first (AddToStore), trains once indexing_size docs exist, then feeds the index
  for (int i = 0; i < first; i++) raw->Add(i, &xb[(size_t)i * d]);
  if (cpu->Indexing()) {
    printf("{\"error\":\"cpu Indexing failed\"}\n");
    return 5;
  }
  {  // same trained state for both (what Dump/Load would give): copy quantizers into the plugin's embedded CPU model
    GammaIVFPQIndex *c = dynamic_cast<GammaIVFPQIndex *>(cpu), *g = dynamic_cast<GammaIVFPQIndex *>(gpu);
    faiss::IndexFlat *cf = dynamic_cast<faiss::IndexFlat *>(c->quantizer);
    g->quantizer->reset();
    g->quantizer->add(c->nlist, cf->xb.data());
    g->quantizer->is_trained = true;
    g->pq.centroids = c->pq.centroids;
    g->is_trained = true;
    g->use_precomputed_table = 0;
    g->precompute_table();
  }
  if (gpu->Indexing()) {
    printf("{\"error\":\"B200IVFPQ Indexing failed: %s\"}\n", gb200_last_error());
    return 5;
  }
  auto feed = [&](int from, int to) {
    for (int s = from; s < to; s += 1000) {  // AddRTVecsToIndex chunks (vector_manager.cc:280-382)
      int n = std::min(1000, to - s);
      ScopeVectors h;
      std::vector<int> lens;
      raw->GetVectorHeader(s, n, h, lens);
      int off = 0;
      for (size_t j = 0; j < h.Size(); j++) {
        for (RetrievalModel *m : {cpu, gpu, cpu_flat, gpu_flat})
          if (!m->Add(lens[j], h.Get(j))) {
            printf("{\"error\":\"Add failed: %s\"}\n", gb200_last_error());
            exit(6);
          }
        off += lens[j];
      }
    }
  };
  feed(0, first);

  PerfTool perf;
  auto search = [&](RetrievalModel *m, const char *rjson, bool has_rank, MultiRangeQueryResults *mr, float lo, float hi,
                    std::vector<float> &D, std::vector<int64_t> &I) {
    GammaSearchCondition cond(&perf);
    cond.topn = k;
    cond.has_rank = has_rank;
    cond.range_query_result = mr;
    cond.Init(lo, hi, bm, raw);
    cond.retrieval_params_ = m->Parse(rjson);
    D.assign((size_t)nq * k, 0.f);
    I.assign((size_t)nq * k, -1);
    return m->Search(&cond, nq, (const uint8_t *)xq.data(), k, D.data(), I.data());
  };
  const float FMAX = std::numeric_limits<float>::max();
  std::vector<float> D0, D1;
  std::vector<int64_t> I0, I1;
  int fails = 0;
  auto report = [&](const char *name, double min_same, double max_rel) {
    Cmp c = compare(nq, k, D0.data(), I0.data(), D1.data(), I1.data());
    double same = (double)c.same_id / c.slots;
    bool ok = same >= min_same && c.max_rel <= max_rel && c.filled_mismatch == 0;
    printf("{\"case\":\"%s\",\"ids_identical\":%.5f,\"max_rel_err_where_same\":%.3g,\"filled_mismatch\":%ld,\"ok\":%s}\n", name,
           same, c.max_rel, c.filled_mismatch, ok ? "true" : "false");
    if (!ok) fails++;
  };
  const char *rj = "{\"nprobe\":16,\"recall_num\":100,\"metric_type\":\"L2\"}";
  search(cpu, rj, true, nullptr, -FMAX, FMAX, D0, I0);
  search(gpu, rj, true, nullptr, -FMAX, FMAX, D1, I1);
  report("ivfpq_rerank", 0.995, 1e-6);
  search(cpu, rj, false, nullptr, -FMAX, FMAX, D0, I0);
  search(gpu, rj, false, nullptr, -FMAX, FMAX, D1, I1);
  report("ivfpq_adc", 0.98, 1e-4);

  // realtime: add the rest, delete some docs, update one, filter by a range bitmap
  for (int i = first; i < N; i++) raw->Add(i, &xb[(size_t)i * d]);
  feed(first, N);
  std::vector<int64_t> dele;
  for (int i = 0; i < N; i += 97) dele.push_back(i);
  for (int64_t v : dele) bm->Set((uint32_t)v);
  for (RetrievalModel *m : {cpu, gpu, cpu_flat, gpu_flat}) m->Delete(dele);
  {
    std::vector<int64_t> ids(1, 4242);
    std::vector<const uint8_t *> vecs(1, (const uint8_t *)&xb[(size_t)(N - 1) * d]);
    raw->UpdateToStore(4242, (uint8_t *)vecs[0], d * sizeof(float));
    for (RetrievalModel *m : {cpu, gpu, cpu_flat, gpu_flat}) m->Update(ids, vecs);
  }
  MultiRangeQueryResults mr;
  {
    RangeQueryResult r;
    r.SetRange(1000, N - 777);
    r.Resize();
    int cnt = 0;
    for (int doc = 1000; doc <= N - 777; doc++)
      if ((doc * 2654435761u >> 7) % 10 < 3) {
        r.Set(doc - r.MinAligned());
        cnt++;
      }
    r.SetDocNum(cnt);
    mr.Add(std::move(r));
  }
  search(cpu, rj, true, &mr, -FMAX, FMAX, D0, I0);
  search(gpu, rj, true, &mr, -FMAX, FMAX, D1, I1);
  report("ivfpq_rerank_filter_delete_update", 0.995, 1e-6);
  const char *fj = "{\"metric_type\":\"L2\",\"parallel_on_queries\":0}";
  search(cpu_flat, fj, true, &mr, -FMAX, FMAX, D0, I0);
  search(gpu_flat, fj, true, &mr, -FMAX, FMAX, D1, I1);
  report("flat_filter_delete_update", 1.0, 0.0);
  float lo = D0[3], hi = D0[7];
  search(cpu_flat, fj, true, nullptr, lo, hi, D0, I0);
  search(gpu_flat, fj, true, nullptr, lo, hi, D1, I1);
  report("flat_score_window", 1.0, 0.0);
  // ---- docs deleted through the bitmap alone (DelDocByQuery sets bits without calling RetrievalModel::Delete,
  // search/gamma_engine.cc:866): the reference tests the bitmap live, the plugin re-syncs it before every search
  search(cpu, rj, true, nullptr, -FMAX, FMAX, D0, I0);
  {
    int hidden = 0;
    for (int q = 0; q < nq && hidden < 40; q += 3)
      if (I0[(size_t)q * k] >= 0 && !bm->Test((uint32_t)I0[(size_t)q * k])) {
        bm->Set((uint32_t)I0[(size_t)q * k]);
        hidden++;
      }
  }
  search(cpu, rj, true, nullptr, -FMAX, FMAX, D0, I0);
  search(gpu, rj, true, nullptr, -FMAX, FMAX, D1, I1);
  report("ivfpq_bitmap_only_delete", 0.995, 1e-6);
  search(cpu_flat, fj, true, nullptr, -FMAX, FMAX, D0, I0);
  search(gpu_flat, fj, true, nullptr, -FMAX, FMAX, D1, I1);
  report("flat_bitmap_only_delete", 1.0, 0.0);

  // ---- Update + CompactBucket on the host (RealTimeMemData::CompactIfNeed, realtime_mem_data.cc:354-424): delete 40 % of
  // the largest bucket, then update 24 docs; the plugin replaces the touched / compacted lists on the device
  {
    GammaIVFPQIndex *c = dynamic_cast<GammaIVFPQIndex *>(cpu);
    size_t best = 0, best_len = 0;
    for (size_t l = 0; l < c->nlist; l++) {
      long *ids = nullptr;
      size_t len = 0;
      uint8_t *cds = nullptr;
      if (c->rt_invert_index_ptr_->GetIvtList(l, ids, len, cds) && len > best_len) best = l, best_len = len;
    }
    long *ids = nullptr;
    size_t len = 0;
    uint8_t *cds = nullptr;
    c->rt_invert_index_ptr_->GetIvtList(best, ids, len, cds);
    std::vector<int64_t> dele2;
    for (size_t j = 0; j < len && dele2.size() < len * 2 / 5; j++)
      if (!(ids[j] & realtime::kDelIdxMask) && !bm->Test((uint32_t)ids[j])) dele2.push_back(ids[j]);
    for (int64_t v : dele2) bm->Set((uint32_t)v);
    for (RetrievalModel *m : {cpu, gpu, cpu_flat, gpu_flat}) m->Delete(dele2);
    const long compacted_before = c->rt_invert_index_ptr_->cur_ptr_->cur_invert_ptr_->compacted_num_;
    std::vector<int64_t> uids;
    std::vector<const uint8_t *> uvecs;
    for (int t = 0; t < 24; t++) {
      int64_t id = 5000 + 311 * t;
      if (bm->Test((uint32_t)id)) continue;
      const float *nv = &xb[(size_t)((id * 7 + 13) % N) * d];
      raw->UpdateToStore((int)id, (uint8_t *)nv, d * sizeof(float));
      uids.push_back(id);
      uvecs.push_back((const uint8_t *)nv);
    }
    for (RetrievalModel *m : {cpu, gpu, cpu_flat, gpu_flat})
      if (m->Update(uids, uvecs)) {
        printf("{\"error\":\"Update failed: %s\"}\n", gb200_last_error());
        return 7;
      }
    const long compacted = c->rt_invert_index_ptr_->cur_ptr_->cur_invert_ptr_->compacted_num_ - compacted_before;
    printf("{\"host_compacted_postings\":%ld,\"updated\":%zu}\n", compacted, uids.size());
    if (compacted <= 0) fails++;  // the case must exercise CompactBucket
  }
  search(cpu, rj, true, nullptr, -FMAX, FMAX, D0, I0);
  search(gpu, rj, true, nullptr, -FMAX, FMAX, D1, I1);
  report("ivfpq_rerank_after_update_and_compaction", 0.995, 1e-6);
  search(cpu, rj, false, &mr, -FMAX, FMAX, D0, I0);
  search(gpu, rj, false, &mr, -FMAX, FMAX, D1, I1);
  report("ivfpq_adc_filter_after_compaction", 0.98, 1e-4);

  // ---- Dump / Load in the reference's own file format (gamma_index_ivfpq.cc:958-1048, index/gamma_index_io.cc:140-196):
  // a dump written by the reference IVFPQ model is loaded by a fresh B200IVFPQ, and the other way round
  {
    utils::make_dir("/tmp/b200_plugin_parity/dump_cpu");
    utils::make_dir("/tmp/b200_plugin_parity/dump_b200");
    if (cpu->Dump("/tmp/b200_plugin_parity/dump_cpu") || gpu->Dump("/tmp/b200_plugin_parity/dump_b200")) {
      printf("{\"error\":\"Dump failed\"}\n");
      return 8;
    }
    RetrievalModel *gpu2 = reflector().GetNewModel("B200IVFPQ");
    RetrievalModel *cpu2 = reflector().GetNewModel("IVFPQ");
    gpu2->vector_ = raw;
    cpu2->vector_ = raw;
    if (gpu2->Init(json, first) || cpu2->Init(json, first)) {
      printf("{\"error\":\"Init of the loading models failed: %s\"}\n", gb200_last_error());
      return 8;
    }
    int n_loaded = gpu2->Load("/tmp/b200_plugin_parity/dump_cpu");
    int n_loaded_cpu = cpu2->Load("/tmp/b200_plugin_parity/dump_b200");
    printf("{\"loaded_into_b200\":%d,\"loaded_into_reference\":%d}\n", n_loaded, n_loaded_cpu);
    if (n_loaded <= 0 || n_loaded_cpu <= 0) fails++;
    search(cpu, rj, true, nullptr, -FMAX, FMAX, D0, I0);
    search(gpu2, rj, true, nullptr, -FMAX, FMAX, D1, I1);
    report("b200_loads_reference_dump", 0.995, 1e-6);
    search(cpu2, rj, false, nullptr, -FMAX, FMAX, D0, I0);
    search(gpu, rj, false, nullptr, -FMAX, FMAX, D1, I1);
    report("reference_loads_b200_dump", 0.98, 1e-4);
    search(gpu, rj, false, nullptr, -FMAX, FMAX, D0, I0);
    search(gpu2, rj, false, nullptr, -FMAX, FMAX, D1, I1);
    report("loaded_b200_equals_live_b200", 1.0, 0.0);
    delete gpu2;
    delete cpu2;
  }

  // ---- IVFFLAT: the reference model (over this MemoryRawVector: check_vector_ off, as its faiss-like facade does) against
  // B200IVFFLAT from the same reflector — train, Add, filter + deletions, Update, 8 concurrent searches
  {
    struct RefIVFFlat : GammaIndexIVFFlat {
      RefIVFFlat() { check_vector_ = false; }
    };
    RetrievalModel *cpu_if = new RefIVFFlat();
    RetrievalModel *gpu_if = reflector().GetNewModel("B200IVFFLAT");
    if (!gpu_if) {
      printf("{\"error\":\"reflector did not return B200IVFFLAT\"}\n");
      return 9;
    }
    cpu_if->vector_ = raw;
    gpu_if->vector_ = raw;
    const char *ij = "{\"ncentroids\":128,\"metric_type\":\"L2\",\"nprobe\":12}";
    if (cpu_if->Init(ij, N) || gpu_if->Init(ij, N)) {
      printf("{\"error\":\"IVFFLAT Init failed: %s\"}\n", gb200_last_error());
      return 9;
    }
    if (cpu_if->Indexing()) {
      printf("{\"error\":\"IVFFLAT Indexing failed\"}\n");
      return 9;
    }
    {  // same trained quantizer on both sides
      GammaIndexIVFFlat *c = dynamic_cast<GammaIndexIVFFlat *>(cpu_if), *g = dynamic_cast<GammaIndexIVFFlat *>(gpu_if);
      faiss::IndexFlat *cf = dynamic_cast<faiss::IndexFlat *>(c->quantizer);
      g->quantizer->reset();
      g->quantizer->add(c->nlist, cf->xb.data());
      g->quantizer->is_trained = true;
      g->is_trained = true;
    }
    for (int s0 = 0; s0 < N; s0 += 1000) {
      int nn = std::min(1000, N - s0);
      ScopeVectors h;
      std::vector<int> lens;
      raw->GetVectorHeader(s0, nn, h, lens);
      for (size_t j = 0; j < h.Size(); j++)
        for (RetrievalModel *m : {cpu_if, gpu_if})
          if (!m->Add(lens[j], h.Get(j))) {
            printf("{\"error\":\"IVFFLAT Add failed: %s\"}\n", gb200_last_error());
            return 9;
          }
    }
    const char *irj = "{\"nprobe\":12,\"metric_type\":\"L2\"}";
    search(cpu_if, irj, false, nullptr, -FMAX, FMAX, D0, I0);
    search(gpu_if, irj, false, nullptr, -FMAX, FMAX, D1, I1);
    report("ivfflat", 0.995, 0.0);
    search(cpu_if, irj, false, &mr, -FMAX, FMAX, D0, I0);
    search(gpu_if, irj, false, &mr, -FMAX, FMAX, D1, I1);
    report("ivfflat_filter_deleted", 0.995, 0.0);
    {
      std::vector<int64_t> uids;
      std::vector<const uint8_t *> uvecs;
      for (int t = 0; t < 16; t++) {
        int64_t id = 7000 + 173 * t;
        if (bm->Test((uint32_t)id)) continue;
        const float *nv = &xb[(size_t)((id * 11 + 5) % N) * d];
        raw->UpdateToStore((int)id, (uint8_t *)nv, d * sizeof(float));
        uids.push_back(id);
        uvecs.push_back((const uint8_t *)nv);
      }
      for (RetrievalModel *m : {cpu_if, gpu_if, cpu, gpu, cpu_flat, gpu_flat})
        if (m->Update(uids, uvecs)) {
          printf("{\"error\":\"IVFFLAT Update failed: %s\"}\n", gb200_last_error());
          return 9;
        }
    }
    search(cpu_if, irj, false, nullptr, -FMAX, FMAX, D0, I0);
    search(gpu_if, irj, false, nullptr, -FMAX, FMAX, D1, I1);
    report("ivfflat_after_update", 0.995, 0.0);
  }

  // ---- concurrent Search (the engine's normal mode, tests/test.h:1033-1062): 8 threads, results equal the serial ones
  {
    search(gpu, rj, true, nullptr, -FMAX, FMAX, D0, I0);
    int bad = 0;
#pragma omp parallel for num_threads(8) reduction(+ : bad)
    for (int t = 0; t < 32; t++) {
      std::vector<float> Dt;
      std::vector<int64_t> It;
      PerfTool pt;
      GammaSearchCondition cond(&pt);
      cond.topn = k;
      cond.has_rank = true;
      cond.range_query_result = nullptr;
      cond.Init(-FMAX, FMAX, bm, raw);
      cond.retrieval_params_ = gpu->Parse(rj);
      Dt.assign((size_t)nq * k, 0.f);
      It.assign((size_t)nq * k, -1);
      int rc = gpu->Search(&cond, nq, (const uint8_t *)xq.data(), k, Dt.data(), It.data());
      if (rc || memcmp(It.data(), I0.data(), It.size() * sizeof(int64_t)) || memcmp(Dt.data(), D0.data(), Dt.size() * sizeof(float)))
        bad++;
    }
    printf("{\"case\":\"concurrent_search_8_threads\",\"mismatching_calls\":%d,\"ok\":%s}\n", bad, bad ? "false" : "true");
    if (bad) fails++;
  }
  printf("{\"plugin_parity\":\"%s\",\"gpu_mem_bytes\":%ld}\n", fails ? "FAIL" : "PASS", gpu->GetTotalMemBytes());
  return fails ? 1 : 0;
}
