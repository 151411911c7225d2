// Multi-GPU exchange behind the C-ABI (reference analogue: faiss IndexReplicas / IndexShards inside
// index/impl/gpu/gamma_gpu_cloner.cpp:209-212, which merge on the host).  One process per GPU, the index replicated, the
// batch sharded by query: after its own search every rank PUSHES its [n][k] result straight into every peer's result
// window with plain stores over NVLink (peer memory mapped through CUDA IPC) and raises an epoch flag there; the same
// kernel then waits for the peers' flags.  One launch, no host round trip, no collective library on the data path.
//
//   window (per rank, double buffered by epoch parity):  [2][world][slot_bytes] results + [2][world] u32 flags
//   exchange kernel, CTA p: copy my slot -> peer p's window (16-byte stores), __threadfence_system, flag[p's view of me]
//                           = epoch (release, system scope); then spin on MY flag from peer p (acquire, system scope).
// Ordering across calls: a rank can run at most one exchange ahead of a peer (it needs the peer's flag of the
// previous epoch to finish), and the peer raises that flag only after its stream has passed the consumers of the epoch
// before — two buffers are enough as long as each rank consumes results in stream order.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/gamma_b200.h"
#include "kernels.h"

namespace gb {

__global__ void __launch_bounds__(256) exchange_push_wait_kernel(const uint4 *__restrict__ mine, uint4 *const *peer_slot,
                                                                 uint32_t *const *peer_flag, const uint32_t *my_flags,
                                                                 int rank, int world, long long n16, uint32_t epoch,
                                                                 unsigned int *err) {
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
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + p) : "memory");
        if (clock64() - t0 > 8000000000LL) {  // ~4 s: a peer never arrived (it failed or was never called) — do not hang
          atomicExch(err, 1u + (unsigned)p);
          break;
        }
      } while ((int32_t)(v - epoch) < 0);
    }
  }
}

}  // namespace gb

struct gb200_comm {
  int device = 0, rank = 0, world = 1;
  long long slot_bytes = 0;  // one rank's result: [n*k] f32 + [n*k] i64, rounded up to 16
  unsigned char *win = nullptr;       // [2][world][slot_bytes] then [2][world] u32 flags
  std::vector<unsigned char *> peer;  // mapped windows of every rank (own pointer at [rank])
  uint4 **d_peer_slot = nullptr;      // [2][world] where MY slot lives in every peer's window
  uint32_t **d_peer_flag = nullptr;   // [2][world] MY flag in every peer's window
  unsigned int *d_err = nullptr;      // set by the exchange kernel when a peer did not arrive in time
  uint32_t epoch = 0;
  bool connected = false;
  size_t win_bytes() const { return (size_t)2 * world * slot_bytes + (size_t)2 * world * sizeof(uint32_t); }
  unsigned char *slot_of(unsigned char *w, int buf, int r) const { return w + ((size_t)buf * world + r) * slot_bytes; }
  uint32_t *flags_of(unsigned char *w, int buf) const {
    return reinterpret_cast<uint32_t *>(w + (size_t)2 * world * slot_bytes) + (size_t)buf * world;
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
  std::vector<uint4 *> slots((size_t)2 * c->world);
  std::vector<uint32_t *> flags((size_t)2 * c->world);
  for (int b = 0; b < 2; b++)
    for (int p = 0; p < c->world; p++) {
      slots[(size_t)b * c->world + p] = reinterpret_cast<uint4 *>(c->slot_of(c->peer[p], b, c->rank));
      flags[(size_t)b * c->world + p] = c->flags_of(c->peer[p], b) + c->rank;
    }
  if (cudaMalloc(&c->d_peer_slot, slots.size() * sizeof(uint4 *)) != cudaSuccess ||
      cudaMalloc(&c->d_peer_flag, flags.size() * sizeof(uint32_t *)) != cudaSuccess ||
      cudaMalloc(&c->d_err, sizeof(unsigned int)) != cudaSuccess)
    return GB200_ENOMEM;
  cudaMemset(c->d_err, 0, sizeof(unsigned int));
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
  if (c->win) cudaFree(c->win);
  delete c;
  return GB200_OK;
}

// where this rank's search writes its own result (D at +0, I at +n*k*4, as bench.py's packed buffer) and where the
// gathered result of the NEXT exchange will be readable
int gb200_comm_buffers(gb200_comm *c, void **my_slot, void **all_slots) {
  if (!c) return GB200_EINVAL;
  const int buf = (int)((c->epoch + 1) & 1);
  if (my_slot) *my_slot = c->slot_of(c->win, buf, c->rank);
  if (all_slots) *all_slots = c->slot_of(c->win, buf, 0);
  return GB200_OK;
}

// push the slot written since the last exchange to every peer and wait for theirs; everything on `stream`
int gb200_comm_exchange(gb200_comm *c, int64_t bytes, void *stream) {
  if (!c || !c->connected || bytes <= 0 || bytes > c->slot_bytes) return GB200_EINVAL;
  if (cudaSetDevice(c->device) != cudaSuccess) return GB200_ECUDA;
  c->epoch++;
  const int buf = (int)(c->epoch & 1);
  if (c->world > 1) {
    gb::exchange_push_wait_kernel<<<c->world, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4 *>(c->slot_of(c->win, buf, c->rank)), c->d_peer_slot + (size_t)buf * c->world,
        c->d_peer_flag + (size_t)buf * c->world, c->flags_of(c->win, buf), c->rank, c->world, (bytes + 15) / 16, c->epoch,
        c->d_err);
    if (cudaGetLastError() != cudaSuccess) return GB200_ECUDA;
  }
  return GB200_OK;
}

int gb200_ivfpq_search_sharded(gb200_index *ix, gb200_comm *c, int n, const float *xq_dev, int k,
                               const gb200_search_params *sp, float **D_all, int64_t **I_all_of_rank0, void *stream) {
  if (!ix || !c || n <= 0 || k <= 0) return GB200_EINVAL;
  const long long per = (long long)n * k * 12;
  if (per > c->slot_bytes) return GB200_EINVAL;
  void *mine = nullptr, *all = nullptr;
  gb200_comm_buffers(c, &mine, &all);
  float *D = static_cast<float *>(mine);
  int64_t *I = reinterpret_cast<int64_t *>(static_cast<unsigned char *>(mine) + (size_t)n * k * 4);
  int rc = gb200_ivfpq_search_dev(ix, n, xq_dev, k, sp, D, I, stream);
  if (rc != GB200_OK) return rc;
  rc = gb200_comm_exchange(c, per, stream);
  if (rc != GB200_OK) return rc;
  // rank r's block: [n*k] f32 distances then [n*k] i64 ids at all + r * slot_bytes
  if (D_all) *D_all = static_cast<float *>(all);
  if (I_all_of_rank0) *I_all_of_rank0 = reinterpret_cast<int64_t *>(static_cast<unsigned char *>(all) + (size_t)n * k * 4);
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
