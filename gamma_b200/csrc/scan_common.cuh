// Device helpers shared by the IVFPQ scan kernels (ivfpq_scan.cu, ivfpq_scan_v3.cu): mbarrier / TMA bulk-copy
// wrappers, raw shared-memory lookups, L2 prefetch, the per-probe record.
#pragma once
#include "common.cuh"

// dynamic shared memory at file scope so PTX can name it: its shared-window address is a link-time
// constant that ptxas folds into the LDS immediate (no per-lookup base add).
extern __shared__ __align__(16) unsigned char gb_scan_smem[];

// Dynamic shared memory starts right after the bytes the driver reserves per CTA (cudaDevAttrReservedSharedMemoryPerBlock,
// 1 KB on sm_90+), so "prmt + const + 4*s" is the complete shared-window address of a table word.  The host checks the
// attribute at index creation and the kernels trap if the assumption does not hold.
#define GB_SMEM_RESERVED 1024

namespace gb {

struct ProbeInfo {
  long long off;  // first posting of the list in the pools
  int len;        // postings visible to the scan (retrieve_idx_pos_)
  int rank;       // probe rank in the query's coarse ordering (tie-break order)
  float dis0;
};

// Extent of one inverted list as the scan may use it: len (retrieve_idx_pos_) is loaded with acquire semantics, off
// after it.  The writer publishes off first and len with release semantics (publish_lists_kernel), so a new len is
// never paired with an old off while appends / relocations / compactions run next to searches
// (RealTimeMemData::ExtendBucketMem's copy-swap, realtime/realtime_mem_data.cc:426-474).
__device__ __forceinline__ void load_list_extent(const long long *list_off, const int *list_len, int key, long long &off,
                                                 int &len) {
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(len) : "l"(list_len + key) : "memory");
  asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(off) : "l"(list_off + key) : "memory");
}

// ---- mbarrier / TMA bulk-copy wrappers (cp.async.bulk -> UBLKCP in SASS)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                             unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void l2_prefetch_line(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ uint32_t prmt_v(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
// table word at shared-window address off + GB_SMEM_RESERVED + 4 * S0 (immediate)
template <int S0>
__device__ __forceinline__ float lds_raw(uint32_t off) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(off), "n"(GB_SMEM_RESERVED + 4 * S0));
  return v;
}

// packed fp32 pairs (sm_100: add.f32x2 -> FADD2, one issue slot for two adds)
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

}  // namespace gb
