// K3 — per-query merge of the scan's survivors, optional exact re-rank on the raw vectors,
// score window and final top-k.  Replaces compute_dis (index/impl/gamma_index_ivfpq.cc:642-697):
//   has_rank : for every recall candidate dis = fvec_L2sqr / fvec_inner_product(xi, raw[vid], raw_d),
//              keep if min_score <= dis <= max_score, k-heap, heap_reorder          (:646-680)
//   !has_rank: heap_reorder(recall_num); copy the first k entries passing the window (:681-696)
// and the vid lookup the scan deferred (ids[list_off + pos]).
//
// exact_distance() reproduces the summation order of faiss' AVX kernels
// (faiss utils/distances_simd.cpp:366-431): 8 strided partial sums (mul then add, unfused),
// hi/lo halves added, optional 4-wide and masked tails (fused), then two horizontal adds —
// so re-ranked distances are bit-identical to the CPU engine built with the same flags.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace gb {

constexpr int RR_THREADS = 128;

template <bool IP>
__global__ void __launch_bounds__(RR_THREADS) rerank_kernel(RerankParams P, int p2_all, int p2_r) {
  extern __shared__ __align__(16) unsigned char smem[];
  u64 *keys = reinterpret_cast<u64 *>(smem);                      // [p2_all]
  u64 *keys2 = keys + p2_all;                                     // [p2_r]
  int *vids = reinterpret_cast<int *>(keys2 + p2_r);              // [p2_r]
  float *qs = reinterpret_cast<float *>(vids + p2_r);             // [raw_d]
  const int q = blockIdx.x, tid = threadIdx.x;
  // rows of this query in the [S][R] candidate block; sort only what is there
  const int total = (P.nsplit ? min(P.nsplit[q], P.S) : (q < P.n_full ? 1 : P.S)) * P.R;
  const u64 *cand = P.cand + (size_t)q * P.S * P.R;
  p2_all = next_pow2(total);
  for (int i = tid; i < p2_all; i += RR_THREADS) keys[i] = i < total ? cand[i] : GB_KEY_MAX;
  if (P.has_rank)
    for (int i = tid; i < P.raw_d; i += RR_THREADS) qs[i] = P.xq[(size_t)q * P.xq_stride + i];
  __syncthreads();
  block_bitonic_sort(keys, p2_all);  // ascending: best candidates first, scan order among equals

  // recall set = first R finite keys; resolve vids
  for (int i = tid; i < p2_r; i += RR_THREADS) {
    int vid = -1;
    if (i < P.R) {
      u64 k = keys[i];
      if (k != GB_KEY_MAX) {
        uint32_t seq = (uint32_t)k;
        int rank = seq >> GB_SEQ_POS_BITS, pos = seq & GB_SEQ_POS_MASK;
        int list = P.keys[(size_t)q * P.nprobe + rank];
        vid = P.ids[P.list_off[list] + pos];
      }
    }
    vids[i] = vid;
  }
  __syncthreads();
  if (P.has_rank && P.stage_rows <= 0) {  // unstaged path: pull the candidates' raw rows towards L2 now
    const int lines = (P.raw_d * 4 + 127) >> 7;
    const int nr = min(P.R, p2_r);
    for (int i = tid; i < nr * lines; i += RR_THREADS) {
      const int c = i / lines, l = i - c * lines;
      const int vid = vids[c];
      if (vid >= 0 && (long long)vid < P.nraw)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.raw + (size_t)vid * P.raw_d + l * 32));
    }
  }

  float *od = P.out_dist + (size_t)q * P.k;
  long long *oi = P.out_ids + (size_t)q * P.k;
  const float neutral = IP ? -3.402823466e38f : 3.402823466e38f;

  if (P.has_rank) {
    // exact distances, 8 lanes per candidate
    const int sub = tid & 7, oct = tid >> 3;
    const int n_oct = RR_THREADS / 8;
    if (P.stage_rows > 0) {
      // The candidates' raw rows (random 4 * raw_d-byte rows of a multi-GB table) are pulled into shared memory with
      // cp.async, every row of a chunk requested before the first one is used: hundreds of DRAM lines in flight per CTA
      // instead of the handful the L2-prefetch + load-on-use scheme sustained (profiles/r02: 1.4 TB/s, long-scoreboard
      // stall 8.5 per issue).  Rows are padded by 8 floats so the four octets of a warp read different banks.
      const int pitch = P.raw_d + 8;
      float *rows = reinterpret_cast<float *>(smem + P.stage_off);
      const int nr = min(P.R, p2_r), lane = tid & 31, warp = tid >> 5;
      for (int i = tid; i < p2_r; i += RR_THREADS) keys2[i] = GB_KEY_MAX;
      for (int c0 = 0; c0 < nr; c0 += P.stage_rows) {
        const int ncur = min(P.stage_rows, nr - c0);
        for (int c = warp; c < ncur; c += RR_THREADS / 32) {
          const int vid = vids[c0 + c];
          if (vid >= 0 && (long long)vid < P.nraw) {
            const float *src = P.raw + (size_t)vid * P.raw_d;
            for (int off = lane * 4; off < P.raw_d; off += 128)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(
                               rows + (size_t)c * pitch + off)),
                           "l"(src + off)
                           : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        for (int base = 0; base < ncur; base += n_oct) {
          const int c = base + oct, i = c0 + c;
          const int vid = c < ncur ? vids[i] : -1;
          const bool have = vid >= 0 && (long long)vid < P.nraw;
          const float dis = exact_distance_octet<IP, true>(qs, rows + (size_t)(have ? c : 0) * pitch, have ? P.raw_d : 0, sub);
          if (sub == 0 && c < ncur) {
            const bool ok = have && dis <= P.max_score && dis >= P.min_score;  // IsSimilarScoreValid
            keys2[i] = ok ? (((u64)dist_to_key32<IP>(dis) << 32) | (uint32_t)i) : GB_KEY_MAX;
          }
        }
        __syncthreads();  // the next chunk overwrites the rows
      }
    } else {
      for (int base = 0; base < p2_r; base += n_oct) {
        int i = base + oct;
        int vid = i < p2_r ? vids[i] : -1;
        bool have = vid >= 0 && (long long)vid < P.nraw;
        const float *y = P.raw + (size_t)(have ? vid : 0) * P.raw_d;
        float dis = exact_distance_octet<IP>(qs, y, have ? P.raw_d : 0, sub);
        if (sub == 0 && i < p2_r) {
          bool ok = have && dis <= P.max_score && dis >= P.min_score;  // IsSimilarScoreValid
          keys2[i] = ok ? (((u64)dist_to_key32<IP>(dis) << 32) | (uint32_t)i) : GB_KEY_MAX;
        }
      }
    }
    __syncthreads();
    block_bitonic_sort(keys2, p2_r);
    for (int j = tid; j < P.k; j += RR_THREADS) {
      u64 k = j < p2_r ? keys2[j] : GB_KEY_MAX;
      float dv = neutral;
      long long iv = -1;
      if (k != GB_KEY_MAX) {
        dv = key32_to_dist<IP>((uint32_t)(k >> 32));
        iv = vids[(uint32_t)k];
      }
      od[j] = dv;
      oi[j] = iv;
      for (int p = 0; p < P.sink.n_peers; p++) {  // multi-GPU: the same row into every peer's window (NVLink stores)
        P.sink.dist[p][(size_t)q * P.k + j] = dv;
        P.sink.ids[p][(size_t)q * P.k + j] = iv;
      }
    }
  } else {
    // ADC distances are final: window filter in sorted order, first k
    if (tid < 32) {
      int filled = 0;
      for (int base = 0; base < p2_r && filled < P.k; base += 32) {
        int i = base + tid;
        bool ok = false;
        float dis = 0.f;
        if (i < p2_r && vids[i] >= 0) {
          dis = key32_to_dist<IP>((uint32_t)(keys[i] >> 32));
          ok = dis <= P.max_score && dis >= P.min_score;
        }
        unsigned m = __ballot_sync(GB_FULL, ok);
        int slot = filled + __popc(m & ((1u << tid) - 1u));
        if (ok && slot < P.k) {
          od[slot] = dis;
          oi[slot] = vids[i];
        }
        filled += __popc(m);
      }
      if (filled > P.k) filled = P.k;
      for (int j = filled + tid; j < P.k; j += 32) {
        od[j] = neutral;
        oi[j] = -1;
      }
      if (P.sink.n_peers) {  // multi-GPU: copy the finished row (this warp wrote all of it)
        __syncwarp();
        for (int j = tid; j < P.k; j += 32) {
          const float dv = od[j];
          const long long iv = oi[j];
          for (int p = 0; p < P.sink.n_peers; p++) {
            P.sink.dist[p][(size_t)q * P.k + j] = dv;
            P.sink.ids[p][(size_t)q * P.k + j] = iv;
          }
        }
      }
    }
  }
  if (P.sink.n_peers) {
    // The CTA's stores into the peers' windows are ordered before thread 0's system-scope fence by the barrier (the
    // pattern of a collective's "store, barrier, one thread fences and signals"); the CTA then counts itself done.  The
    // last CTA of the grid has (through the counter) observed every other CTA's fence, so the flags it raises with
    // release stores become visible at the peers after all rows of this rank.  It also does this exchange's wait: for
    // the peers' flags of wait_epoch (deferred form: the previous exchange, normally long there).
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();
      const unsigned int prev = atomicAdd(P.sink.done, 1u);
      if (prev == gridDim.x - 1) {
        *P.sink.done = 0u;  // for the next launch (stream order)
        __threadfence_system();
        for (int p = 0; p < P.sink.n_peers; p++)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(P.sink.flag[p]), "r"(P.sink.epoch) : "memory");
        for (int p = 0; p < P.sink.n_peers; p++) {
          if (P.sink.wait[p] == nullptr) continue;
          uint32_t v;
          const long long t0 = clock64();
          do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(P.sink.wait[p]) : "memory");
            if (clock64() - t0 > 8000000000LL) {  // ~4 s: the peer failed or was never called — do not hang
              atomicExch(P.sink.err, 1u + (unsigned)P.sink.peer_rank[p]);
              break;
            }
          } while ((int32_t)(v - P.sink.wait_epoch) < 0);
        }
      }
    }
  }
}

cudaError_t launch_rerank(const RerankParams &P_in, cudaStream_t st) {
  RerankParams P = P_in;
  int p2_all = next_pow2(P.S * P.R);
  int p2_r = next_pow2(P.R);
  size_t smem = (size_t)(p2_all + p2_r) * sizeof(u64) + (size_t)p2_r * sizeof(int) + (size_t)P.raw_d * sizeof(float);
  // row staging for the exact re-rank: as many 16-candidate rounds as fit in ~64 KB
  P.stage_rows = 0;
  P.stage_off = 0;
  if (P.has_rank && (P.raw_d & 3) == 0 && !P.no_stage) {
    const size_t row_bytes = (size_t)(P.raw_d + 8) * sizeof(float);
    int rows = (int)((64 * 1024) / row_bytes) & ~15;
    const int want = (std::min(P.R, p2_r) + 15) & ~15;
    if (rows > want) rows = want;
    if (rows >= 16) {
      P.stage_off = (int)((smem + 15) & ~(size_t)15);
      P.stage_rows = rows;
      smem = (size_t)P.stage_off + (size_t)rows * row_bytes;
    }
  }
  {
    cudaError_t e = P.is_ip ? ensure_dynamic_smem(rerank_kernel<true>, smem) : ensure_dynamic_smem(rerank_kernel<false>, smem);
    if (e != cudaSuccess) return e;
  }
  if (P.is_ip)
    rerank_kernel<true><<<P.n, RR_THREADS, smem, st>>>(P, p2_all, p2_r);
  else
    rerank_kernel<false><<<P.n, RR_THREADS, smem, st>>>(P, p2_all, p2_r);
  return cudaGetLastError();
}

}  // namespace gb
