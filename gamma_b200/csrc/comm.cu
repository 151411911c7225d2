// Multi-GPU exchange behind the C-ABI (reference analogue: faiss IndexReplicas / IndexShards inside
// index/impl/gpu/gamma_gpu_cloner.cpp:209-212, which merge on the host).  One process per GPU, the index replicated, the
// batch sharded by query: after its own search every rank PUSHES its [n][k] result straight into every peer's result
// window with plain stores over NVLink (peer memory mapped through CUDA IPC) and raises an epoch flag there; the same
// kernel then waits for the peers' flags.  One launch, no host round trip, no collective library on the data path.
//
//   window (per rank, four buffers, buffer = epoch % 4):  [4][world][slot_bytes] results + [4][world] u32 flags
//   exchange kernel, CTA p: copy my slot -> peer p's window (16-byte stores), __threadfence_system, flag[p's view of me]
//                           = epoch (release, system scope); then spin on MY flag from peer p (acquire, system scope).
// Fused form (what gb200_ivfpq_search_sharded* normally run): there is no exchange kernel at all — the search's final
// kernel (rerank_kernel, one CTA per query) stores each query's k results into every peer's window right where it
// writes them locally, every CTA fences at system scope and counts itself done, and the last CTA raises the flags and
// does the wait (kernels.h PeerSink, rerank.cu).  The separate kernel below serves gb200_comm_exchange and searches that
// end in another kernel.
// Two ways to wait:
//   * gb200_ivfpq_search_sharded waits for the peers' flags of THIS epoch: the gathered result of this call is readable
//     when the stream gets past the kernel;
//   * gb200_ivfpq_search_sharded_deferred waits for the flags of the PREVIOUS epoch and hands back that epoch's window: a
//     pipelined server consumes the gathered results one call late, and no rank ever idles for the slowest rank of the
//     current step (its result travels while the next search runs).  gb200_comm_flush waits for the last epoch.
// Ordering across calls (why four buffers): a peer pushes epoch e + 4 into the buffer my stream reads epoch e from.  Its
// exchange kernel pushes BEFORE it waits, so that push only follows the peer's exchange of epoch e + 3, which (deferred
// mode, the weaker one) waited for MY flag of epoch e + 2.  My stream raises that flag in the exchange of epoch e + 2, and
// my consumers of epoch e sit between my exchanges e + 1 and e + 2 (deferred) or e and e + 1 (immediate) — before the
// flag either way, as long as each rank consumes results in stream order.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/gamma_b200.h"
#include "kernels.h"

namespace gb {

__global__ void __launch_bounds__(256) exchange_push_wait_kernel(const uint4 *__restrict__ mine, uint4 *const *peer_slot,
                                                                 uint32_t *const *peer_flag, const uint32_t *wait_flags,
                                                                 int rank, int world, long long n16, uint32_t epoch,
                                                                 uint32_t wait_epoch, unsigned int *err) {
  const int p = blockIdx.x;  // one CTA per peer
  if (p != rank) {
    uint4 *dst = peer_slot[p];
    for (long long i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = mine[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag[p]), "r"(epoch) : "memory");
      uint32_t v;
      const long long t0 = clock64();
      if (wait_flags == nullptr) return;  // deferred wait, first call: nothing to wait for yet
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(wait_flags + p) : "memory");
        if (clock64() - t0 > 8000000000LL) {  // ~4 s: a peer never arrived (it failed or was never called) — do not hang
          atomicExch(err, 1u + (unsigned)p);
          break;
        }
      } while ((int32_t)(v - wait_epoch) < 0);
    }
  }
}

// wait for every peer's flag of `wait_epoch` (gb200_comm_flush)
__global__ void exchange_wait_kernel(const uint32_t *wait_flags, int rank, int world, uint32_t wait_epoch, unsigned int *err) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= world || p == rank) return;
  uint32_t v;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(wait_flags + p) : "memory");
    if (clock64() - t0 > 8000000000LL) {
      atomicExch(err, 1u + (unsigned)p);
      break;
    }
  } while ((int32_t)(v - wait_epoch) < 0);
}

}  // namespace gb

struct gb200_comm {
  int device = 0, rank = 0, world = 1;
  long long slot_bytes = 0;  // one rank's result: [n*k] f32 + [n*k] i64, rounded up to 16
  unsigned char *win = nullptr;       // [NBUF][world][slot_bytes] then [NBUF][world] u32 flags
  std::vector<unsigned char *> peer;  // mapped windows of every rank (own pointer at [rank])
  uint4 **d_peer_slot = nullptr;      // [NBUF][world] where MY slot lives in every peer's window
  uint32_t **d_peer_flag = nullptr;   // [NBUF][world] MY flag in every peer's window
  unsigned int *d_err = nullptr;      // set by the exchange kernel when a peer did not arrive in time
  unsigned int *d_done = nullptr;     // CTA counter of the fused form (the search's final kernel feeds the peers itself)
  uint32_t epoch = 0;
  bool connected = false;
  static constexpr int NBUF = 4;
  size_t win_bytes() const { return (size_t)NBUF * world * slot_bytes + (size_t)NBUF * world * sizeof(uint32_t); }
  unsigned char *slot_of(unsigned char *w, int buf, int r) const { return w + ((size_t)buf * world + r) * slot_bytes; }
  uint32_t *flags_of(unsigned char *w, int buf) const {
    return reinterpret_cast<uint32_t *>(w + (size_t)NBUF * world * slot_bytes) + (size_t)buf * world;
  }
};

extern "C" {

int gb200_comm_create(int device, int rank, int world, int64_t slot_bytes, gb200_comm **out, uint8_t *handle) {
  if (!out || !handle || world < 1 || rank < 0 || rank >= world || slot_bytes <= 0) return GB200_EINVAL;
  if (sizeof(cudaIpcMemHandle_t) > GB200_COMM_HANDLE_BYTES) return GB200_EUNSUPPORTED;
  if (cudaSetDevice(device) != cudaSuccess) return GB200_ECUDA;
  gb200_comm *c = new gb200_comm;
  c->device = device, c->rank = rank, c->world = world;
  c->slot_bytes = (slot_bytes + 15) & ~15LL;
  if (cudaMalloc(&c->win, c->win_bytes()) != cudaSuccess || cudaMemset(c->win, 0, c->win_bytes()) != cudaSuccess) {
    delete c;
    return GB200_ENOMEM;
  }
  memset(handle, 0, GB200_COMM_HANDLE_BYTES);
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, c->win) != cudaSuccess) {
    cudaFree(c->win);
    delete c;
    return GB200_ECUDA;
  }
  memcpy(handle, &h, sizeof(h));
  cudaDeviceSynchronize();
  *out = c;
  return GB200_OK;
}

int gb200_comm_connect(gb200_comm *c, const uint8_t *handles) {
  if (!c || !handles || c->connected) return GB200_EINVAL;
  if (cudaSetDevice(c->device) != cudaSuccess) return GB200_ECUDA;
  c->peer.assign(c->world, nullptr);
  for (int p = 0; p < c->world; p++) {
    if (p == c->rank) {
      c->peer[p] = c->win;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)p * GB200_COMM_HANDLE_BYTES, sizeof(h));
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      return GB200_ECUDA;
    }
    c->peer[p] = static_cast<unsigned char *>(ptr);
  }
  std::vector<uint4 *> slots((size_t)gb200_comm::NBUF * c->world);
  std::vector<uint32_t *> flags((size_t)gb200_comm::NBUF * c->world);
  for (int b = 0; b < gb200_comm::NBUF; b++)
    for (int p = 0; p < c->world; p++) {
      slots[(size_t)b * c->world + p] = reinterpret_cast<uint4 *>(c->slot_of(c->peer[p], b, c->rank));
      flags[(size_t)b * c->world + p] = c->flags_of(c->peer[p], b) + c->rank;
    }
  if (cudaMalloc(&c->d_peer_slot, slots.size() * sizeof(uint4 *)) != cudaSuccess ||
      cudaMalloc(&c->d_peer_flag, flags.size() * sizeof(uint32_t *)) != cudaSuccess ||
      cudaMalloc(&c->d_err, sizeof(unsigned int)) != cudaSuccess || cudaMalloc(&c->d_done, sizeof(unsigned int)) != cudaSuccess)
    return GB200_ENOMEM;
  cudaMemset(c->d_err, 0, sizeof(unsigned int));
  cudaMemset(c->d_done, 0, sizeof(unsigned int));
  cudaMemcpy(c->d_peer_slot, slots.data(), slots.size() * sizeof(uint4 *), cudaMemcpyHostToDevice);
  cudaMemcpy(c->d_peer_flag, flags.data(), flags.size() * sizeof(uint32_t *), cudaMemcpyHostToDevice);
  c->connected = true;
  return GB200_OK;
}

int gb200_comm_destroy(gb200_comm *c) {
  if (!c) return GB200_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < (int)c->peer.size(); p++)
    if (p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
  if (c->d_peer_slot) cudaFree(c->d_peer_slot);
  if (c->d_peer_flag) cudaFree(c->d_peer_flag);
  if (c->d_err) cudaFree(c->d_err);
  if (c->d_done) cudaFree(c->d_done);
  if (c->win) cudaFree(c->win);
  delete c;
  return GB200_OK;
}

// where this rank's search writes its own result (D at +0, I at +n*k*4, as bench.py's packed buffer) and where the
// gathered result of the NEXT exchange will be readable
int gb200_comm_buffers(gb200_comm *c, void **my_slot, void **all_slots) {
  if (!c) return GB200_EINVAL;
  const int buf = (int)((c->epoch + 1) % gb200_comm::NBUF);
  if (my_slot) *my_slot = c->slot_of(c->win, buf, c->rank);
  if (all_slots) *all_slots = c->slot_of(c->win, buf, 0);
  return GB200_OK;
}

// push the slot written since the last exchange to every peer, then wait: for this epoch's flags, or (deferred) for the
// previous epoch's; everything on `stream`
static int comm_exchange(gb200_comm *c, int64_t bytes, void *stream, bool deferred) {
  if (!c || !c->connected || bytes <= 0 || bytes > c->slot_bytes) return GB200_EINVAL;
  if (cudaSetDevice(c->device) != cudaSuccess) return GB200_ECUDA;
  c->epoch++;
  const int buf = (int)(c->epoch % gb200_comm::NBUF);
  if (c->world > 1) {
    const uint32_t wait_epoch = deferred ? c->epoch - 1 : c->epoch;
    const uint32_t *wait_flags = wait_epoch == 0 ? nullptr : c->flags_of(c->win, (int)(wait_epoch % gb200_comm::NBUF));
    gb::exchange_push_wait_kernel<<<c->world, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4 *>(c->slot_of(c->win, buf, c->rank)), c->d_peer_slot + (size_t)buf * c->world,
        c->d_peer_flag + (size_t)buf * c->world, wait_flags, c->rank, c->world, (bytes + 15) / 16, c->epoch, wait_epoch,
        c->d_err);
    if (cudaGetLastError() != cudaSuccess) return GB200_ECUDA;
  }
  return GB200_OK;
}

int gb200_comm_exchange(gb200_comm *c, int64_t bytes, void *stream) { return comm_exchange(c, bytes, stream, false); }

static int search_sharded(gb200_index *ix, gb200_comm *c, int n, const float *xq_dev, int k, const gb200_search_params *sp,
                          float **D_all, int64_t **I_all_of_rank0, void *stream, bool deferred) {
  if (!ix || !c || n <= 0 || k <= 0) return GB200_EINVAL;
  const long long per = (long long)n * k * 12;
  if (per > c->slot_bytes) return GB200_EINVAL;
  void *mine = nullptr, *all = nullptr;
  gb200_comm_buffers(c, &mine, &all);
  float *D = static_cast<float *>(mine);
  int64_t *I = reinterpret_cast<int64_t *>(static_cast<unsigned char *>(mine) + (size_t)n * k * 4);
  // Fused form: the search's final kernel stores each query's k results into every peer's window as it produces them,
  // its last CTA raises this rank's flags and does the wait — the exchange costs no launch and no second pass over the
  // results.  (Searches that end in another kernel, or more peers than the sink holds: the separate exchange kernel.)
  gb::PeerSink sk;
  memset(&sk, 0, sizeof(sk));
  const uint32_t e = c->epoch + 1, wait_epoch = deferred ? e - 1 : e;
  const int buf = (int)(e % gb200_comm::NBUF);
  const bool fuse = c->connected && c->world > 1 && c->world - 1 <= gb::GB_MAX_PEERS && ((long long)n * k) % 2 == 0;
  if (fuse) {
    for (int p = 0; p < c->world; p++) {
      if (p == c->rank) continue;
      unsigned char *slot = c->slot_of(c->peer[p], buf, c->rank);
      const int i = sk.n_peers++;
      sk.dist[i] = reinterpret_cast<float *>(slot);
      sk.ids[i] = reinterpret_cast<long long *>(slot + (size_t)n * k * 4);
      sk.flag[i] = c->flags_of(c->peer[p], buf) + c->rank;
      sk.wait[i] = wait_epoch ? c->flags_of(c->win, (int)(wait_epoch % gb200_comm::NBUF)) + p : nullptr;
      sk.peer_rank[i] = p;
    }
    sk.epoch = e;
    sk.wait_epoch = wait_epoch;
    sk.done = c->d_done;
    sk.err = c->d_err;
  }
  int used = 0;
  int rc = gb_ivfpq_search_dev_sink(ix, n, xq_dev, k, sp, D, I, stream, fuse ? &sk : nullptr, &used);
  if (rc != GB200_OK) return rc;
  if (used) {
    c->epoch = e;  // pushed, flagged and awaited by the search's final kernel
  } else {
    rc = comm_exchange(c, per, stream, deferred);
    if (rc != GB200_OK) return rc;
  }
  if (deferred)  // the window of the epoch before the one just pushed (nullptr on the first call)
    all = c->epoch >= 2 ? c->slot_of(c->win, (int)((c->epoch - 1) % gb200_comm::NBUF), 0) : nullptr;
  // rank r's block: [n*k] f32 distances then [n*k] i64 ids at all + r * slot_bytes
  if (D_all) *D_all = static_cast<float *>(all);
  if (I_all_of_rank0)
    *I_all_of_rank0 = all ? reinterpret_cast<int64_t *>(static_cast<unsigned char *>(all) + (size_t)n * k * 4) : nullptr;
  return GB200_OK;
}

int gb200_ivfpq_search_sharded(gb200_index *ix, gb200_comm *c, int n, const float *xq_dev, int k,
                               const gb200_search_params *sp, float **D_all, int64_t **I_all_of_rank0, void *stream) {
  return search_sharded(ix, c, n, xq_dev, k, sp, D_all, I_all_of_rank0, stream, false);
}

int gb200_ivfpq_search_sharded_deferred(gb200_index *ix, gb200_comm *c, int n, const float *xq_dev, int k,
                                        const gb200_search_params *sp, float **D_all_prev, int64_t **I_all_prev_of_rank0,
                                        void *stream) {
  return search_sharded(ix, c, n, xq_dev, k, sp, D_all_prev, I_all_prev_of_rank0, stream, true);
}

// wait (on `stream`) until every peer's result of the LAST exchange has arrived; *all_slots = that epoch's window
int gb200_comm_flush(gb200_comm *c, void **all_slots, void *stream) {
  if (!c || !c->connected) return GB200_EINVAL;
  if (cudaSetDevice(c->device) != cudaSuccess) return GB200_ECUDA;
  if (all_slots) *all_slots = c->epoch ? c->slot_of(c->win, (int)(c->epoch % gb200_comm::NBUF), 0) : nullptr;
  if (c->world > 1 && c->epoch) {
    gb::exchange_wait_kernel<<<(c->world + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
        c->flags_of(c->win, (int)(c->epoch % gb200_comm::NBUF)), c->rank, c->world, c->epoch, c->d_err);
    if (cudaGetLastError() != cudaSuccess) return GB200_ECUDA;
  }
  return GB200_OK;
}

int64_t gb200_comm_slot_bytes(gb200_comm *c) { return c ? c->slot_bytes : 0; }

// device -> host copy of (part of) the gathered window on `stream`; sync != 0 also waits for it
int gb200_comm_read(gb200_comm *c, void *dst_host, const void *src_dev, int64_t bytes, void *stream, int sync) {
  if (!c || !dst_host || !src_dev || bytes < 0) return GB200_EINVAL;
  if (cudaSetDevice(c->device) != cudaSuccess) return GB200_ECUDA;
  if (cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess)
    return GB200_ECUDA;
  if (sync && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return GB200_ECUDA;
  return GB200_OK;
}

// 0 = every exchange so far saw all peers; 1 + p = peer p did not arrive within the time limit (synchronises the device)
int gb200_comm_status(gb200_comm *c) {
  if (!c || !c->d_err) return 0;
  cudaSetDevice(c->device);
  unsigned int e = 0;
  if (cudaMemcpy(&e, c->d_err, sizeof(e), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)e;
}

}  // extern "C"
