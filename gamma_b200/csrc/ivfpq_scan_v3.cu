// K2 (v3) — the IVFPQ ADC scan for M = 32 as a PERSISTENT kernel with dynamic work distribution.
// Same arithmetic, table layout, posting layout and selection (BlockTopR) as ivfpq_scan.cu; what changed, and why
// (profiles/r01c_scan_m32_v2_ncu_summary.txt: issue slots 67.7 % busy, 211 issued instructions per 32-posting block,
// 17 % of warp time at the CTA's final barrier, 12 % of the launch is tail):
//
//  * one CTA per resident slot (SMs x CTAs/SM) instead of one CTA per (query, split).  A CTA takes the next query
//    from a global counter, loads its table (TMA bulk copy, one mbarrier reused with alternating parity) and scans;
//    no wave quantisation, no host-side plan.
//  * inside a query the unit of work is an ITEM = up to ch_blocks consecutive 32-posting blocks of ONE probed list.
//    Every warp claims items one at a time with an atomicAdd on the query's counter (the claim for the next item is
//    in flight while the current one is scanned), so warps of a CTA finish within one item of each other — there is
//    no static split to be unlucky with;
//  * the counter lives in global memory, so several CTAs can work on ONE query: when the query queue is empty a
//    CTA that would otherwise idle looks for the running query with the most unclaimed items, takes a candidate row
//    of it (rows[q]), loads the same table and claims from the same counter.  The tail of the launch is split
//    exactly as far as the idle CTAs allow; K3 merges the rows.  (Batches smaller than the machine are spread the
//    same way from the first cycle.)
//  * the loop itself: all per-block addressing derives from ONE incrementing block index (three IMAD.WIDE against
//    per-lane constant bases; existence of the lane's posting, L2-prefetch bound and scan-order word are compares /
//    adds against per-item constants), and the 32 table words are accumulated with packed adds
//    (add.f32x2 -> FADD2): ~115 issued instructions per block instead of ~165.
//
// Reference being replaced (file:line): GammaIVFPQScanner::scan_list_with_table index/impl/gamma_index_ivfpq.h:576-601,
// KnnSearchResults::add :351-370, scan_one_list / the probe loop index/impl/gamma_index_ivfpq.cc:597-640, 790-818,
// RTInvertIndex::GetIvtList realtime/realtime_invert_index.cc:77-81.
#include "scan_common.cuh"

namespace gb {

// shared-memory carve-up (host mirrors it in scan_v3_smem_bytes).  Everything the hot loop touches sits at a
// COMPILE-TIME shared-window address (table at GB_SMEM_RESERVED, control words right behind it), so those accesses need
// no address registers; only the candidate buffer and the item prefix depend on run-time sizes.
//   [lut 64 KB][misc 96 ints][mbar 16 B][ring mbarriers 512 B][ProbeInfo x nprobe][item prefix x (nprobe + 1), padded to 16][buf u64 cap]
//   [posting ring: warps x RING slots x 1280 B]
// misc: [0..1] tau, [2] cnt, [3] tau_f, [4..67] scratch, [68..70] round flags, [72] q, [73] row
constexpr int V3_MISC_OFF = 65536;
constexpr int V3_MBAR_OFF = V3_MISC_OFF + 96 * 4;
constexpr int V3_RBAR_OFF = V3_MBAR_OFF + 16;          // ring-slot mbarriers (TMA-fed ring): 16 warps x 4 slots
constexpr int V3_PINFO_OFF = V3_RBAR_OFF + 16 * 4 * 8;
constexpr int V3_SLOT_BYTES = 1280;  // one 32-posting block in the ring: codes 1024 B, ids 128 B, t(p) 128 B
struct V3Smem {
  unsigned char *ring;
  u64 *buf;
  ProbeInfo *pinfo;
  int *item_prefix;
  uint2 *itab;
  int *misc;
  unsigned long long *mbar;
  unsigned long long *rbar;  // [warp][RING] one mbarrier per ring slot (TMA-fed ring)
};

// per-query probe block: [ProbeInfo x nprobe][item prefix x (nprobe + 1)], padded to 16 B, then the ITEM TABLE
// [max_items x {first pool block, blocks << 16 | probe}] (8 B per item, item ids in claim order)
__host__ __device__ inline size_t v3_itab_off(int nprobe) {
  size_t b = (size_t)nprobe * sizeof(ProbeInfo) + (size_t)(nprobe + 1) * sizeof(int);
  return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t v3_probe_bytes(int nprobe, int max_items) {
  return v3_itab_off(nprobe) + (((size_t)max_items * 8 + 15) & ~(size_t)15);
}
size_t scan_v3_probe_bytes(int nprobe, int max_items) { return v3_probe_bytes(nprobe, max_items); }
size_t scan_v3_smem_bytes(int nprobe, int max_items, int cap, int warps, int ring) {
  return V3_PINFO_OFF + v3_probe_bytes(nprobe, max_items) + (size_t)cap * sizeof(u64) + (size_t)warps * ring * V3_SLOT_BYTES;
}

__device__ __forceinline__ V3Smem v3_carve(unsigned char *smem, int nprobe, int max_items, int cap) {
  V3Smem S;
  S.misc = reinterpret_cast<int *>(smem + V3_MISC_OFF);
  S.mbar = reinterpret_cast<unsigned long long *>(smem + V3_MBAR_OFF);
  S.rbar = reinterpret_cast<unsigned long long *>(smem + V3_RBAR_OFF);
  S.pinfo = reinterpret_cast<ProbeInfo *>(smem + V3_PINFO_OFF);
  S.item_prefix = reinterpret_cast<int *>(smem + V3_PINFO_OFF + (size_t)nprobe * sizeof(ProbeInfo));
  S.itab = reinterpret_cast<uint2 *>(smem + V3_PINFO_OFF + v3_itab_off(nprobe));
  S.buf = reinterpret_cast<u64 *>(smem + V3_PINFO_OFF + v3_probe_bytes(nprobe, max_items));
  S.ring = smem + V3_PINFO_OFF + v3_probe_bytes(nprobe, max_items) + (size_t)cap * sizeof(u64);
  return S;
}

// volatile reads of a control word at a compile-time offset from a shared-window address held in a register
template <int OFF>
__device__ __forceinline__ int lds_ctl_s32(uint32_t base) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1+%2];" : "=r"(v) : "r"(base), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ float lds_ctl_f32(uint32_t base) {
  float v;
  asm volatile("ld.volatile.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(base), "n"(OFF));
  return v;
}
// two adjacent control words with one load (8-byte aligned offset)
template <int OFF>
__device__ __forceinline__ void lds_ctl_s32_f32(uint32_t base, int &a, float &b) {
  static_assert(OFF % 8 == 0, "v2 load");
  asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2+%3];" : "=r"(a), "=f"(b) : "r"(base), "n"(OFF));
}
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}
// base + a * B as one IMAD.WIDE (a: block index, B: bytes per block)
template <int B>
__device__ __forceinline__ const unsigned char *wide_at(const void *base, uint32_t a) {
  const unsigned char *r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "n"(B), "l"(base));
  return r;
}
__device__ __forceinline__ const unsigned char *wide_at_r(const void *base, uint32_t a, uint32_t b) {
  const unsigned char *r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(base));
  return r;
}
// keep a kernel-lifetime per-lane constant in registers (stops the compiler from re-deriving it inside the loop)
template <typename T>
__device__ __forceinline__ T *pin_ptr(T *p) {
  asm volatile("" : "+l"(p));
  return p;
}

// ---- cp.async (LDGSTS): asynchronous global -> shared copies of this lane's own bytes, tracked per thread
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- TMA-fed ring: one elected lane issues bulk copies of whole 32-posting blocks, completion on the slot's mbarrier
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_a(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAITR_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONER_%=;\n\tbra WAITR_%=;\n\tDONER_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ int ld_volatile_s32(const int *p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// K2b (v3) — per-query probe table in the scan's shared-memory layout: list extents, dis0, exclusive prefix of the
// per-list ITEM counts (ceil(blocks / ch_blocks)).  One warp per query.  scan_one_list's list lookup
// (gamma_index_ivfpq.cc:597-640) and dis0 of precompute_list_tables (gamma_index_ivfpq.h:216-230, 236-299).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) probe_setup_v3_kernel(ScanParams P) {
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= P.n) return;
  const int np = P.nprobe, ch = P.ch_blocks;
  unsigned char *dst = P.probe_g + (size_t)q * v3_probe_bytes(np, P.v3_max_items);
  ProbeInfo *pinfo = reinterpret_cast<ProbeInfo *>(dst);
  int *prefix = reinterpret_cast<int *>(dst + (size_t)np * sizeof(ProbeInfo));
  uint2 *itab = reinterpret_cast<uint2 *>(dst + v3_itab_off(np));
  const float *xq = P.xq + (size_t)q * P.d;
  int carry = 0;
  unsigned my_postings = 0, n_full = 0;
  for (int j0 = 0; j0 < np; j0 += 32) {
    const int j = j0 + lane;
    ProbeInfo pi;
    pi.off = 0, pi.len = 0, pi.rank = j, pi.dis0 = 0.f;
    if (j < np) {
      const int key = P.keys[(size_t)q * np + j];
      if (key >= 0 && key < P.nlist) {  // scan_one_list: key < 0 or >= nlist => skip (gamma_index_ivfpq.cc:602-609)
        load_list_extent(P.list_off, P.list_len, key, pi.off, pi.len);
        if (P.is_ip) {  // dis0 = <q, centroid>
          const float *cen = P.centroids + (size_t)key * P.d;
          float s = 0.f;
          for (int i = 0; i < P.d; i++) s = fmaf(__ldg(xq + i), __ldg(cen + i), s);
          pi.dis0 = s;
        } else {
          pi.dis0 = P.coarse_dis[(size_t)q * np + j];
        }
      }
      pinfo[j] = pi;
    }
    const int nb = (pi.len + 31) >> 5;
    const int ni = (nb + ch - 1) / ch;
    int incl = ni;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(GB_FULL, incl, o);
      if (lane >= o) incl += v;
    }
    if (j < np) prefix[j] = carry + incl - ni;
    carry += __shfl_sync(GB_FULL, incl, 31);
    n_full += __reduce_add_sync(GB_FULL, (unsigned)(nb / ch));
    my_postings += (unsigned)pi.len;
  }
  // ---- item table.  Claim order = table order: all FULL items (ch blocks) of every list first, the lists' shorter
  // remainders last, so that the warps of a CTA run out of work within a short item of each other (the scan-order key
  // of a posting does not depend on who scans it when).  If the table cannot hold every item (a list outgrew the host's
  // bound) the order is list-major, the one the scan's search fallback assumes for the items beyond the table.
  const bool reorder = carry <= P.v3_max_items;
  int c_full = 0, c_rem = 0, c_all = 0;
  for (int j0 = 0; j0 < np; j0 += 32) {
    const int j = j0 + lane;
    int nb = 0;
    uint32_t offb = 0;
    if (j < np) {
      nb = (pinfo[j].len + 31) >> 5;  // written by this very lane above
      offb = (uint32_t)(pinfo[j].off >> 5);
    }
    const int nf = nb / ch, rem = nb - nf * ch, ni = nf + (rem ? 1 : 0);
    int i_full = nf, i_rem = rem ? 1 : 0, i_all = ni;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(GB_FULL, i_full, o), b = __shfl_up_sync(GB_FULL, i_rem, o),
                c = __shfl_up_sync(GB_FULL, i_all, o);
      if (lane >= o) i_full += a, i_rem += b, i_all += c;
    }
    if (reorder) {
      int it = c_full + i_full - nf;
      for (int k = 0; k < nf; k++, it++) itab[it] = make_uint2(offb + (uint32_t)(k * ch), ((uint32_t)ch << 16) | (uint32_t)j);
      if (rem) itab[(int)n_full + c_rem + i_rem - 1] = make_uint2(offb + (uint32_t)(nf * ch), ((uint32_t)rem << 16) | (uint32_t)j);
    } else {
      int it = c_all + i_all - ni;
      for (int b0 = 0; b0 < nb && it < P.v3_max_items; b0 += ch, it++)
        itab[it] = make_uint2(offb + (uint32_t)b0, ((uint32_t)min(ch, nb - b0) << 16) | (uint32_t)j);
    }
    c_full += __shfl_sync(GB_FULL, i_full, 31);
    c_rem += __shfl_sync(GB_FULL, i_rem, 31);
    c_all += __shfl_sync(GB_FULL, i_all, 31);
  }
  my_postings = __reduce_add_sync(GB_FULL, my_postings);
  if (lane == 0) {
    prefix[np] = carry;  // may exceed the table's capacity (a list grew past the host's bound): those items are located by search
    if (P.scanned) atomicAdd(P.scanned, (unsigned long long)my_postings);
  }
}

cudaError_t launch_probe_setup_v3(const ScanParams &P, cudaStream_t st) {
  probe_setup_v3_kernel<<<(P.n + 7) / 8, 256, 0, st>>>(P);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// the scan of one (query, row) by one CTA
//
// Posting pipeline.  profiles/r02c showed what bounds the register-prefetch + L2-hint scheme of v2 / early v3: the L2's
// sector-lookup rate (88.8 M lookups per launch = 12.6 TB/s, its ceiling), because every posting sector is looked up
// twice — once by the prefetch hint, once by the load — and ~40 % of the loads still miss.  Here every warp keeps RING
// 32-posting blocks in flight as REAL asynchronous copies (cp.async, each lane copies exactly the bytes it will read:
// 2 x 16 B of codes, its id, its t(p)) into a private ring in shared memory, plus one block in registers: one L2 lookup
// per sector, RING + 1 blocks of latency cover per warp, no hints, no cross-lane synchronisation (a lane only ever reads
// what it copied itself, cp.async.wait_group is per thread).  The copy engine runs ahead of the consumer across item
// boundaries: the next item is claimed and located while the first block of the current one is scanned.
// ---------------------------------------------------------------------------------------------------------------
template <bool IP, bool HAS_VALID, int WARPS, int PER, int RING, bool TMA>
__device__ __forceinline__ void scan_loop_m32_v3(const ScanParams &P, const V3Smem &S, BlockTopR &topr, const int q,
                                                 const int it_first, uint32_t &ring_epoch, const uint32_t lut_parity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lane4 = lane * 4;
  const int np = P.nprobe;
  const int n_items = S.item_prefix[np];
  const int ch = P.ch_blocks;
  // lane * 0 (a run-time zero the compiler cannot see through): with a provably warp-uniform address ptxas rewrites the
  // claim below into its warp-aggregated form, whose broadcast shuffle waits for the atomic at the point of issue
  int *const claim = P.v3_claim + q + lane * P.v3_zero;
  const int soft_limit = (int)pin_u32((uint32_t)(P.cap - WARPS * 32));
  // ids at or beyond the bitmap's size (appended after this search began) and negative ids (dead / padding) fail
  const uint32_t valid_lim = (uint32_t)(P.valid_bits < 0x7fffffffLL ? P.valid_bits : 0x7fffffffLL);
  const uint32_t ctl = pin_u32(smem_u32(S.misc));  // shared-window address of the control words
  volatile int *flags = S.misc + 68;  // 3 rotating slots: bit0 = prune wanted, bit1 = work left

  // per-lane bases, constant for the whole kernel: a block is addressed by its index in the pool (bi = posting / 32)
  //   codes: 1 KB per block, this lane's 16 B of chunk 0 at lane * 16, of chunk 1 at 512 + lane * 16
  const unsigned char *const codes_lane = pin_ptr(P.codes + lane * 16);
  const int *const ids_lane = pin_ptr(P.ids + lane);
  const float *const nrm_lane = pin_ptr(P.norms + lane);
  // this lane's bytes of ring slot 0: codes chunk 0 at +0, chunk 1 at +512; id at ring_w + 1024, t(p) at ring_w + 1152
  const uint32_t ring_c = pin_u32(smem_u32(S.ring) + (uint32_t)warp * (RING * V3_SLOT_BYTES) + lane * 16);
  const uint32_t ring_w = pin_u32(smem_u32(S.ring) + (uint32_t)warp * (RING * V3_SLOT_BYTES) + lane * 4);
  // TMA-fed ring: warp-uniform slot base and this warp's slot barriers
  const uint32_t ring_b = smem_u32(S.ring) + (uint32_t)warp * (RING * V3_SLOT_BYTES);
  const uint32_t rbar = smem_u32(S.rbar) + (uint32_t)warp * (4 * 8);

  // ---- consumer's item.  bi / bi_stop / bi_end are warp-uniform; bil, seqc differ per lane
  uint32_t bi = 0, bi_end = 0;  // next block to TAKE from the ring / end of the current item
  uint32_t bi_stop = 0;         // next block index at which the slow path below has something to do
  uint32_t bil = 0;             // this lane's posting of block b exists  <=>  b < bil
  uint32_t seqc = 0;            // scan-order word of this lane's posting in block b = seqc + 32 * (b + 1)
  float dis0 = 0.f;
  // ---- copy engine (producer): runs up to RING blocks ahead.  It walks the same items as the consumer, one segment
  // [pbi, pbi_end) at a time; the item located next waits in [qbi, qbi_end) (empty when qbi == qbi_end) and is popped
  // when the current segment is fully requested — no hand-shake with the consumer
  uint32_t pbi = 0, pbi_end = 0;  // next block to REQUEST / end of the producer's segment
  uint32_t qbi = 0, qbi_end = 0;  // queued segment
  // blocks requested / taken so far (slot = count % RING).  The counts run on across the queries of this CTA: the
  // TMA-fed ring's slot barriers keep their phase between queries (slot use u = count / RING waits with parity u & 1)
  uint32_t pn = ring_epoch, cn = ring_epoch;
  // ---- claims
  int it_next = it_first;       // lane 0: the claim in flight (the first one was issued by the caller)
  bool claim_pending = true;    // a claim has been issued and not consumed yet
  bool nx_valid = false;        // the located item this warp scans next: list nx_j, blocks [nx_bi, nx_bi_end)
  uint32_t nx_j = 0;
  uint32_t nx_bi = 0, nx_bi_end = 0;

  // One claim per item: an atomic add on the query's counter by lane 0.  The result stays in lane 0's register until
  // the slow path consumes it one block later (the aggregated form waited for it at the point of issue —
  // profiles/r02a: long-scoreboard stall 4.9 per issue).
  auto claim_issue = [&]() {
    if (lane == 0) asm volatile("atom.global.add.s32 %0, [%1], 1;" : "=r"(it_next) : "l"(claim) : "memory");
    claim_pending = true;
  };
  // request the producer's next block into ring slot pn % RING (all lanes: each copies its own bytes).  Branch-free for
  // the per-lane cp.async ring: the copies are predicated, a (possibly empty) group is committed either way.
  auto refill_one = [&]() {
    const bool seg_done = pbi == pbi_end;  // pop the queued segment
    pbi = seg_done ? qbi : pbi;
    pbi_end = seg_done ? qbi_end : pbi_end;
    qbi = seg_done ? qbi_end : qbi;
    const bool go = (pbi != pbi_end) && (pn - cn < (uint32_t)RING);
    const uint32_t so = (pn % RING) * V3_SLOT_BYTES;
    if (TMA) {
      // whole block by three bulk copies from one elected lane (no LSU work, no per-lane addresses): the slot's
      // mbarrier flips when all bytes have landed.  Slot use u = pn / RING waits with parity u & 1.
      if (go && elect_one()) {
        const uint32_t bar = rbar + (pn % RING) * 8;
        mbar_expect_tx_a(bar, IP ? 1152u : 1280u);
        tma_bulk_g2s_a(ring_b + so, wide_at<1024>(P.codes, pbi), 1024u, bar);
        tma_bulk_g2s_a(ring_b + so + 1024, wide_at<128>(P.ids, pbi), 128u, bar);
        if (!IP) tma_bulk_g2s_a(ring_b + so + 1152, wide_at<128>(P.norms, pbi), 128u, bar);
      }
    } else {
      const unsigned char *cp = wide_at<1024>(codes_lane, pbi);
      const uint32_t g = go ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "@p cp.async.cg.shared.global [%0], [%2], 16;\n\t"
          "@p cp.async.cg.shared.global [%0+512], [%2+512], 16;\n\t"
          "@p cp.async.ca.shared.global [%1+1024], [%3], 4;\n\t}" ::"r"(ring_c + so),
          "r"(ring_w + so), "l"(cp), "l"(wide_at<128>(ids_lane, pbi)), "r"(g)
          : "memory");
      if (!IP)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0+1152], [%1], 4;\n\t}" ::"r"(
                         ring_w + so),
                     "l"(wide_at<128>(nrm_lane, pbi)), "r"(g)
                     : "memory");
      cp_async_commit();
    }
    pn += go ? 1u : 0u;
    pbi += go ? 1u : 0u;
  };
  // Slow path, run when bi == bi_stop (warp-uniform), i.e. after the FIRST block of an item and at its END:
  //  A. a claim is in flight: take its result and locate that item (list j, blocks) so that the copy engine can run
  //     into it before the consumer gets there;
  //  B. the current item is finished: open the located one and issue the claim for the one after it.
  auto slow_path = [&]() {
    if (claim_pending) {
      claim_pending = false;
      const int it = __shfl_sync(GB_FULL, it_next, 0);
      nx_valid = it < n_items;
      // the query has no unclaimed item left: what remains are the items other warps are inside.  The first warp to
      // notice fetches the CTA's NEXT (query, row) from the queue now — late enough not to reserve work a faster CTA
      // could take, early enough for the two dependent atomics to be back before the query's last barrier
      if (!nx_valid && lane == 0 && S.misc[77] && atomicExch(&S.misc[76], 1) == 0) {
        int qn = atomicAdd(P.v3_next_q, 1);
        int rw = 0;
        if (qn < P.n) rw = atomicAdd(P.v3_rows + qn, 1);
        else qn = -1;
        S.misc[74] = qn;
        S.misc[75] = rw;
      }
      if (nx_valid) {
        if (it < P.v3_max_items) {  // the item table gives list and blocks directly
          const uint2 e = S.itab[it];
          nx_j = e.y & 0xffffu;
          nx_bi = e.x;
          nx_bi_end = e.x + (e.y >> 16);
        } else {  // beyond the table (the host sized it from the longest list it knew): last j with item_prefix[j] <= it
          int j = -1;
#pragma unroll 1
          for (int j0 = 0; j0 < np; j0 += 32) {
            const int v = (j0 + lane < np) ? S.item_prefix[j0 + lane] : 0x7fffffff;
            j += __popc(__ballot_sync(GB_FULL, v <= it));
          }
          const ProbeInfo pi = S.pinfo[j];
          const uint32_t offb = (uint32_t)(pi.off >> 5);
          const uint32_t nblk = (uint32_t)(pi.len + 31) >> 5;
          const uint32_t b0 = (uint32_t)(it - S.item_prefix[j]) * (uint32_t)ch;
          nx_j = (uint32_t)j;
          nx_bi = offb + b0;
          nx_bi_end = offb + min(b0 + (uint32_t)ch, nblk);
        }
        qbi = nx_bi;  // the producer's queue is empty here: it popped the current item before its first block was taken
        qbi_end = nx_bi_end;
      }
    }
    if (bi == bi_end) {
      if (nx_valid) {
        nx_valid = false;
        const ProbeInfo pi = S.pinfo[nx_j];
        const uint32_t offb = (uint32_t)(pi.off >> 5);
        bi = nx_bi;
        bi_end = nx_bi_end;
        bil = offb + ((uint32_t)(pi.len - lane + 31) >> 5);
        seqc = ((uint32_t)pi.rank << GB_SEQ_POS_BITS) + (uint32_t)lane - (offb << 5) - 32u;
        dis0 = pi.dis0;
        claim_issue();
        bi_stop = bi + 1;  // consume that claim after the first block of this item (== bi_end for one-block items)
      }                    // else: the query has no unclaimed item left; bi == bi_end == bi_stop stays
    } else {
      bi_stop = bi_end;
    }
  };

  // ---- the block in registers (per lane): 32 pre-rotated code bytes, vid, t(p)
  uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
  int id_n = -1;
  float nrm_n = 0.f;
  bool have_n = false;
  // take the consumer's next block out of the ring into the registers
  auto take_next = [&]() {
    if (bi == bi_stop) slow_path();  // warp-uniform, twice per item
    have_n = bi != bi_end;
    if (have_n) {
      // ring not full although there is something to request (query start, or the next item was located too late for
      // the producer to run ahead): top it up now
      if (pn - cn < (uint32_t)RING && (pbi != pbi_end || qbi != qbi_end)) {
#pragma unroll 1
        for (int k = 0; k < RING; k++) refill_one();
      }
      if (TMA) {
        mbar_wait_a(rbar + (cn % RING) * 8, (cn / RING) & 1u);
      } else {
        // ring full: the groups of the RING - 1 younger blocks (and possibly empty groups) follow the group of the block
        // being taken, so "all but the RING - 1 most recent groups" covers it; otherwise wait for everything
        if (pn - cn == (uint32_t)RING) cp_async_wait<RING - 1>();
        else cp_async_wait<0>();
      }
      const uint32_t so = (cn % RING) * V3_SLOT_BYTES;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3) : "r"(ring_c + so));
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+512];" : "=r"(c4), "=r"(c5), "=r"(c6), "=r"(c7) : "r"(ring_c + so));
      int idv;
      float nv = 0.f;
      asm volatile("ld.shared.s32 %0, [%1+1024];" : "=r"(idv) : "r"(ring_w + so));
      if (!IP) asm volatile("ld.shared.f32 %0, [%1+1152];" : "=f"(nv) : "r"(ring_w + so));
      const bool ex = bi < bil;  // the copy engine does not look at list ends; postings beyond them are ignored here
      id_n = ex ? idv : -1;
      nrm_n = nv;
      cn++;
      bi++;
    }
  };

  u64 skey = 0;
  bool spend = false;
  auto try_append = [&](bool pass, u64 key) -> bool {
    const unsigned m = __ballot_sync(GB_FULL, pass);
    if (m == 0) return false;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(topr.cnt, __popc(m));
    base = __shfl_sync(GB_FULL, base, leader);
    const int slot = base + __popc(m & ((1u << lane) - 1u));
    bool pending = pass;
    if (pass && slot < topr.cap) {
      topr.buf[slot] = key;
      pending = false;
    }
    spend = pending;
    skey = key;
    return __any_sync(GB_FULL, pending);
  };

  bool stalled = false;
  take_next();  // the first claim of this query was issued before the probe block arrived (claim_pending == true)
  mbar_wait(&S.mbar[1], lut_parity);  // the lookup table, while this warp's first blocks travel
  int round = 0;
  for (;;) {
    if (stalled) stalled = try_append(spend && skey < topr.threshold(), skey);  // after a prune
    bool over = false;
    if (!stalled) {
      while (have_n) {  // warp-uniform
        // ---- table addresses (code << 8 | lane * 4) 16 at a time; 32 conflict-free lookups, accumulated as two
        // packed pairs.  The first group consumes the registers the last take_next() filled, so the ring slot they
        // came from is free: the copy engine re-uses it right here.
        uint32_t a[16];
        u64 s01, s23;
#define GB_ADDR4(W, I)                     \
  a[I + 0] = prmt_v(W, lane4, 0x5504);     \
  a[I + 1] = prmt_v(W, lane4, 0x5514);     \
  a[I + 2] = prmt_v(W, lane4, 0x5524);     \
  a[I + 3] = prmt_v(W, lane4, 0x5534);
#define GB_LOOK4(I, O)                                                          \
  s01 = f2_add(s01, f2_pack(lds_raw<O + 0>(a[I + 0]), lds_raw<O + 1>(a[I + 1]))); \
  s23 = f2_add(s23, f2_pack(lds_raw<O + 2>(a[I + 2]), lds_raw<O + 3>(a[I + 3])));
        GB_ADDR4(c0, 0) GB_ADDR4(c1, 4) GB_ADDR4(c2, 8) GB_ADDR4(c3, 12)
        // what the admission test needs of the block in registers, evaluated NOW (asm volatile pins the order) so that
        // its registers are dead before the next block is taken.
        //   nb = dis0 + t(p), or NaN when the lane has no posting (padding, beyond the list end, moved away): NaN fails
        //   the admission compare below by itself.
        float nb;
        uint32_t seq;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %3, 0;\n\tadd.f32 %0, %1, %2;\n\t@p mov.b32 %0, 0x7fc00000;\n\t}"
            : "=f"(nb)
            : "f"(dis0), "f"(nrm_n), "r"(id_n));
        asm volatile("mad.lo.u32 %0, %1, 32, %2;" : "=r"(seq) : "r"(bi), "r"(seqc));
        uint32_t vw = 0xffffffffu, vbit = 1u;
        if (HAS_VALID) {
          vw = (uint32_t)id_n < valid_lim ? __ldg(P.valid + (id_n >> 5)) : 0u;  // latency hidden by the lookups
          asm volatile("shf.l.wrap.b32 %0, 1, 1, %1;" : "=r"(vbit) : "r"(id_n));  // 1 << (id & 31)
        }
        if (!TMA) refill_one();
        s01 = f2_pack(lds_raw<0>(a[0]), lds_raw<1>(a[1]));
        s23 = f2_pack(lds_raw<2>(a[2]), lds_raw<3>(a[3]));
        GB_LOOK4(4, 4) GB_LOOK4(8, 8) GB_LOOK4(12, 12)
        GB_ADDR4(c4, 0) GB_ADDR4(c5, 4) GB_ADDR4(c6, 8) GB_ADDR4(c7, 12)
        // TMA ring: every register the slot fed (codes, id, t(p)) has been read by now, so the async-proxy write
        // into it cannot pass a pending generic-proxy read
        if (TMA) refill_one();
        // ---- next block out of the ring, straight into the registers just freed
        take_next();
        GB_LOOK4(0, 16) GB_LOOK4(4, 20) GB_LOOK4(8, 24) GB_LOOK4(12, 28)
#undef GB_ADDR4
#undef GB_LOOK4
        // ---- pre-test, admission.  Every warp reads the counter once per block so that all of them notice a
        // wanted prune within one block, whether or not they append anything themselves.  The pre-test is one float
        // compare against the float image of tau's distance word (NaN fails it, as in the reference's heap compare).
        float tau_f;  // misc[3]
        int cnt_now;  // misc[2]: one 8-byte load for both
        lds_ctl_s32_f32<8>(ctl, cnt_now, tau_f);
        float s0, s1, s2, s3;
        f2_unpack(s01, s0, s1);
        f2_unpack(s23, s2, s3);
        const float dis = nb + ((s0 + s1) + (s2 + s3));
        bool pass = IP ? dis >= tau_f : dis <= tau_f;
        if (HAS_VALID) pass = pass && (vw & vbit);
        if (__any_sync(GB_FULL, pass || cnt_now > soft_limit)) {  // rare
          if (__any_sync(GB_FULL, pass)) {
            const u64 key = ((u64)dist_to_key32<IP>(dis) << 32) | seq;
            stalled = try_append(pass && key < topr.threshold(), key);
          }
          // appenders re-read the counter after their own atomicAdd: a warp starts a block only while
          // cnt <= cap - 32 * WARPS, so the buffer cannot overflow between sync points
          over = *((volatile int *)topr.cnt) > soft_limit;
          if (stalled || over) break;
        }
      }
    }
    const bool more = stalled || have_n;
    over = over || stalled;
    const int slot = round % 3;
    if (lane == 0 && (more || over)) atomicOr((int *)&flags[slot], (over ? 1 : 0) | (more ? 2 : 0));
    __syncthreads();
    const int v = flags[slot];
    if (threadIdx.x == 0) flags[(round + 2) % 3] = 0;  // used two sync points from now; nobody touches it before
    round++;
    if (v & 1) {  // in-loop prune: approximate (keeps a few more than R, far fewer barriers) when enabled
      if (P.v3_flags & 4) topr.prune_collective<PER, false>();
      else topr.prune_collective<PER, true>();
    }
    if (!(v & 2)) break;
  }
  if (TMA) {  // nothing is in flight here (a warp requests only blocks it consumes); drain anyway so the phases stay in step
    while (cn != pn) {
      mbar_wait_a(rbar + (cn % RING) * 8, (cn / RING) & 1u);
      cn++;
    }
  }
  ring_epoch = pn;
}

// a CTA without a query looks for the running query with the most unclaimed items among the last help_window queries
// and takes a candidate row of it.  Collective; returns false when nothing worth joining is left.
template <int THREADS>
__device__ __forceinline__ bool v3_pick_victim(const ScanParams &P, const V3Smem &S, int &q_out, int &row_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t pbytes = v3_probe_bytes(P.nprobe, P.v3_max_items);
  const size_t total_off = (size_t)P.nprobe * sizeof(ProbeInfo) + (size_t)P.nprobe * sizeof(int);
  const int w0 = max(0, P.n - min(P.help_window, 65536));
  for (;;) {
    uint32_t best = 0;
    for (int qq = P.n - 1 - tid; qq >= w0; qq -= THREADS) {
      const int tot = *reinterpret_cast<const int *>(P.probe_g + (size_t)qq * pbytes + total_off);
      const int cl = ld_volatile_s32(P.v3_claim + qq);
      const int rw = ld_volatile_s32(P.v3_rows + qq);
      const int rem = rw < P.S ? tot - cl : 0;
      if (rem > 0) best = max(best, ((uint32_t)min(rem, 65535) << 16) | (uint32_t)(qq - w0));
    }
    best = __reduce_max_sync(GB_FULL, best);
    if (lane == 0) S.misc[4 + warp] = (int)best;
    __syncthreads();
    if (tid == 0) {
      uint32_t b = 0;
      for (int w = 0; w < THREADS / 32; w++) b = max(b, (uint32_t)S.misc[4 + w]);
      int q = -1, row = 0;
      if ((int)(b >> 16) >= P.help_min) {
        q = w0 + (int)(b & 0xffffu);
        row = atomicAdd(P.v3_rows + q, 1);
        if (row >= P.S) q = -2;  // lost the race for the last row: look again (the query is excluded now)
      }
      S.misc[72] = q;
      S.misc[73] = row;
    }
    __syncthreads();
    const int q = S.misc[72];
    if (q == -1) return false;
    if (q >= 0) {
      q_out = q;
      row_out = S.misc[73];
      return true;
    }
  }
}

template <bool IP, int THREADS, int MINB, int PER, int RING, bool TMA>
__global__ void __launch_bounds__(THREADS, MINB) ivfpq_scan_m32_v3_kernel(ScanParams P) {
  constexpr int WARPS = THREADS / 32;
  const int tid = threadIdx.x;
  const V3Smem S = v3_carve(gb_scan_smem, P.nprobe, P.v3_max_items, P.cap);
  BlockTopR topr;
  topr.buf = S.buf;
  topr.tau = reinterpret_cast<u64 *>(S.misc);
  topr.cnt = S.misc + 2;
  topr.warp_part = S.misc + 4;
  topr.cap = P.cap;
  topr.R = P.R;
  topr.tau_f = reinterpret_cast<float *>(S.misc + 3);
  topr.is_ip = IP ? 1 : 0;
  if (smem_u32(gb_scan_smem) != GB_SMEM_RESERVED) __trap();  // the LDS immediates assume it (host checks the attribute)
  const uint32_t pbytes = (uint32_t)v3_probe_bytes(P.nprobe, P.v3_max_items);
  if (tid == 0) {
    mbar_init(&S.mbar[0], 1);  // probe block (extents, item table)
    mbar_init(&S.mbar[1], 1);  // the query's lookup table
    if (TMA)
      for (int i = 0; i < 16 * 4; i++) mbar_init(&S.rbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t parity = 0;
  uint32_t ring_epoch = 0;
  bool helper = false;
  bool early = false;  // the tables of the query about to start were requested before the previous query's final select
  // the next (query, candidate row) from the queue: fetched by one thread WHILE the current query is scanned (two
  // dependent global atomics, ~2 us) and left in misc[74..75] for the next iteration
  auto fetch_next = [&]() {
    int qn = atomicAdd(P.v3_next_q, 1);
    int rw = 0;
    if (qn < P.n) rw = atomicAdd(P.v3_rows + qn, 1);
    else qn = -1;
    S.misc[74] = qn;
    S.misc[75] = rw;
  };
  // the query's table [256][64] (lut_build_m32_kernel, L2 resident) and its probe block by bulk copies, one mbarrier each
  auto request_tables = [&](int qq) {
    const char *src = reinterpret_cast<const char *>(P.lut_g) + (size_t)qq * 65536;
    mbar_expect_tx(&S.mbar[0], pbytes);
    tma_bulk_g2s(S.pinfo, P.probe_g + (size_t)qq * pbytes, pbytes, &S.mbar[0]);
    mbar_expect_tx(&S.mbar[1], 65536u);
#pragma unroll
    for (int i = 0; i < 4; i++) tma_bulk_g2s(gb_scan_smem + i * 16384, src + i * 16384, 16384u, &S.mbar[1]);
  };
  bool prefetched = false;  // misc[74..75] already hold the next (query, row): fetched inside the previous scan
  for (;;) {
    // ---- the next (query, candidate row) of this CTA
    if (!helper && !prefetched) {
      if (tid == 0) fetch_next();
      __syncthreads();
    }
    int q = S.misc[74], row = S.misc[75];
    __syncthreads();  // everybody has read the pair before it is replaced
    prefetched = false;
    const bool from_queue = !helper && q >= 0;
    if (!from_queue) {
      helper = true;
      if (!v3_pick_victim<THREADS>(P, S, q, row)) break;
    }
    if (row >= P.S) continue;  // every row of this query is taken: the CTAs holding them scan all of its items
    if (tid == 0) {
      *topr.cnt = 0;
      *topr.tau = GB_KEY_MAX;
      *topr.tau_f = IP ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000);  // everything finite is admitted
      S.misc[68] = S.misc[69] = S.misc[70] = 0;
      S.misc[76] = 0;                                        // nobody has fetched the next query yet
      S.misc[77] = (from_queue && (P.v3_flags & 1)) ? 1 : 0;  // ... and whether a warp of the scan should
      if (!early) request_tables(q);
    }
    __syncthreads();
    // every warp's first claim travels to L2 and back while the tables arrive
    int it_first = 0;
    int *const claim0 = P.v3_claim + q + (tid & 31) * P.v3_zero;  // address formed outside the branch (see scan_loop_m32_v3)
    if ((tid & 31) == 0) asm volatile("atom.global.add.s32 %0, [%1], 1;" : "=r"(it_first) : "l"(claim0) : "memory");
    // the small probe block is enough to claim, locate and request the first postings: the 64 KB table is awaited
    // inside the loop, after the first blocks are in flight
    mbar_wait(&S.mbar[0], parity);
    if (P.valid) scan_loop_m32_v3<IP, true, WARPS, PER, RING, TMA>(P, S, topr, q, it_first, ring_epoch, parity);
    else scan_loop_m32_v3<IP, false, WARPS, PER, RING, TMA>(P, S, topr, q, it_first, ring_epoch, parity);
    parity ^= 1u;
    // ---- every warp is past its last table / probe-block read (the loop's closing barrier).  If the next query is
    // known already (fetched inside the scan) its tables can start streaming in under the final select (optional:
    // measured slower — the 64 KB of bulk-copy writes collide with the select's shared-memory traffic)
    early = false;
    if (from_queue && (P.v3_flags & 1)) {
      prefetched = true;  // every warp ends on a failed claim, so one of them fetched
      const int qn = S.misc[74], rn = S.misc[75];
      early = (P.v3_flags & 2) && qn >= 0 && rn < P.S;
      if (early && tid == 0) request_tables(qn);
    }
    // ---- survivors of this CTA -> cand[q][row][0..R)
    topr.prune_collective<PER>();
    const int n_out = min(*((volatile int *)topr.cnt), P.R);
    u64 *out = P.cand + ((size_t)q * P.S + row) * P.R;
    for (int i = tid; i < P.R; i += THREADS) out[i] = i < n_out ? topr.buf[i] : GB_KEY_MAX;
    __syncthreads();
  }
}

template <bool IP, int T, int MINB, int PER, int RING, bool TMA>
static cudaError_t launch_v3_one(const ScanParams &P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = ensure_dynamic_smem(ivfpq_scan_m32_v3_kernel<IP, T, MINB, PER, RING, TMA>, smem);
  if (e != cudaSuccess) return e;
  ivfpq_scan_m32_v3_kernel<IP, T, MINB, PER, RING, TMA><<<grid, T, smem, st>>>(P);
  return cudaGetLastError();
}
template <int T, int MINB, int PER, int RING>
static cudaError_t launch_v3_shape(const ScanParams &P, int grid, cudaStream_t st) {
  const size_t smem = scan_v3_smem_bytes(P.nprobe, P.v3_max_items, P.cap, T / 32, RING);
  if (P.v3_tma) {
    return P.is_ip ? launch_v3_one<true, T, MINB, PER, RING, true>(P, grid, smem, st)
                   : launch_v3_one<false, T, MINB, PER, RING, true>(P, grid, smem, st);
  }
  return P.is_ip ? launch_v3_one<true, T, MINB, PER, RING, false>(P, grid, smem, st)
                 : launch_v3_one<false, T, MINB, PER, RING, false>(P, grid, smem, st);
}

// CTA shapes (threads, CTAs per SM, ring slots per warp): shared memory per CTA = 64 KB table + candidate buffer + probe
// table + warps x slots x 1280 B, two CTAs per SM:
//   384 x 2, ring 2 (default): 24 warps per SM, 3 blocks in flight per warp (2 in the ring + 1 in registers)
//   320 x 2, ring 3: 20 warps per SM, 4 blocks in flight per warp
//   256 x 2, ring 4: 16 warps per SM, 5 blocks in flight per warp
//   512 x 1, ring 2: recall_num > 512 (candidate buffers of 2048 / 4096 keys, 4 / 8 keys per thread in the select)
int scan_v3_ctas_per_sm(int threads, int cap) { return (threads >= 512 || cap > 4 * threads) ? 1 : 2; }
size_t scan_v3_smem_bytes_for(int nprobe, int max_items, int cap, int threads) {
  const int ring = threads == 320 ? 3 : threads == 256 ? 4 : 2;
  return scan_v3_smem_bytes(nprobe, max_items, cap, threads / 32, ring);
}

cudaError_t launch_ivfpq_scan_v3(const ScanParams &P, int grid, cudaStream_t st) {
  if (P.M != 32 || P.nprobe > 2048 || P.cap > 8 * P.m32_threads || (P.cap > 4 * P.m32_threads && P.m32_threads != 512))
    return cudaErrorInvalidValue;
  switch (P.m32_threads) {
    case 512: return P.cap > 2048 ? launch_v3_shape<512, 1, 8, 2>(P, grid, st) : launch_v3_shape<512, 1, 4, 2>(P, grid, st);
    case 448: return launch_v3_shape<448, 2, 4, 2>(P, grid, st);
    case 416: return launch_v3_shape<416, 2, 4, 2>(P, grid, st);
    case 320: return launch_v3_shape<320, 2, 4, 3>(P, grid, st);
    case 256: return launch_v3_shape<256, 2, 4, 4>(P, grid, st);
    default: return launch_v3_shape<384, 2, 4, 2>(P, grid, st);
  }
}

}  // namespace gb
