// Launch-side helper shared by every kernel that needs more than 48 KB of dynamic shared memory.
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

#include "kernels.h"

namespace gb {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (function, device), not of a launch.  Search calls run
// concurrently on several host threads, and the size a kernel needs depends on the call's parameters (nprobe,
// recall_num, d ...): setting the attribute to "this launch's size" right before each launch lets another thread lower it
// between the set and the launch, which then fails with an invalid-value error.  So the attribute only ever grows:
// the largest size requested so far per (function, device) is remembered, and raised under a lock before the first
// launch that needs more.  A launch never needs more than the value in force, whoever set it.
cudaError_t ensure_dynamic_smem(const void *func, size_t smem) {
  if (smem <= 48 * 1024) return cudaSuccess;
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> granted;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> g(mu);
  size_t &cur = granted[std::make_pair(func, dev)];
  if (smem <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) cur = smem;
  return e;
}

}  // namespace gb
