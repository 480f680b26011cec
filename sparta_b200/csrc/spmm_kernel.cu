// Block-sparse (VBR / BELLPACK) x dense SpMM for sm_100a.
//
// Replaces the reference's per-block cublasGemmEx loop over n_streams
// (src/cuda/cuda_utilities.cpp:146-182), the batched cublasSgemmBatched loop
// (:831-868) and the cuSPARSE / CUTLASS Blocked-ELL calls (:1617,
// src/cuda/cutlass_bellpack_lib.cu:157-200) with ONE persistent kernel.
//
// Formulation (the reference's "inverted" product, cuda_utilities.cpp:675-690):
//   Ct[j, r] += sum_k Bt[j, k] * A_blk[r, k]
// so the fixed dimension (128 columns of B) sits on the tcgen05 M axis and the
// variable block height h sits on the N axis (any multiple of 16 up to 256):
// ragged heights cost no padding beyond 16 rows and no dense grid.
//
//   warp 0   : TMA producer.  Per chunk: one 2-D tensor load of the B row-panel
//              (128 columns x 128 bytes of k, SWIZZLE_128B) and one bulk copy of
//              the chunk's packed A images (already in MMA smem byte order).
//   warp 6   : second TMA producer (kSlots variant): the two warps take alternate
//              chunks, because ONE thread cannot start more than one pipeline stage
//              per ~600 cycles whatever its size (profiles/r2_copy_issue_microbench.txt).
//   warp 1   : single-thread tcgen05.mma issuer; accumulators live in TMEM.
//              Member segments with adjacent accumulators are fused into one
//              MMA (N up to 256) so the B panel is read from smem once per run.
//   warps 2-5: epilogue.  tcgen05.ld -> predicated stores to C; re-zero TMEM.
//
// A super-row (several row segments sharing one TMEM allocation) walks the
// MERGED list of its members' column blocks, so a B panel is fetched once per
// super-row instead of once per nonzero block.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "spmm_kernel.h"

namespace sparta {

// ------------------------------------------------------------------ PTX glue
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Note on scopes: every wait below uses the default (CTA-scope) acquire even when the arrival
// comes from the peer CTA.  What those cross-CTA barriers order is shared memory written by TMA
// and TMEM written by tcgen05.st, both consumed by tcgen05.mma behind tcgen05 fences -- no
// generic-proxy global data.  A cluster-scope acquire makes ptxas emit CCTL.IVALL (an L1
// invalidate behind the long scoreboard) after every successful wait, which showed up as the
// top stall in the first ncu capture of the pair kernel.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
// Relaxed flavour for the per-chunk "my half has landed" relay: the data it announces was written
// by TMA and is read by tcgen05.mma (both async proxy, ordered by the mbarrier itself), so no
// generic-proxy release is needed -- and a cluster-scope release costs a full memory barrier
// (ERRBAR/MEMBAR, ~2000 cycles per chunk in the first pair-mode trace).
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// One lane of a CONVERGED warp.  tcgen05.mma, TMA and bulk-copy instructions take their
// operands in uniform registers; issued under `if (lane == 0)` ptxas wraps each of them in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (~250 cycles per MMA in the first trace).
// Under elect.sync it knows exactly one lane is active and moves the operands over directly.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Spin on an mbarrier phase.  A pipeline bug must not wedge the GPU box, so the
// wait traps after 5 s instead of spinning forever.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int tag) {
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0x3FF) == 0 && global_timer_ns() - t0 > 5000000000ull) {
      printf("sparta spmm: mbarrier wait timed out (cta %d thread %d tag %d parity %u)\n",
             blockIdx.x, threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity, tag);
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes,
                                          uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// Pull a byte range into L2 ahead of the copy that will need it (no shared memory involved).
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// Completion of all previously issued MMAs arrives on `bar`; in pair mode on the barrier at
// the same offset in BOTH CTAs (the peer's producer / epilogue wait on their own copy).
template <bool kPair>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if constexpr (kPair) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
        " [%0], %1;" ::"r"(bar),
        "h"(static_cast<uint16_t>(3))
        : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     bar)
                 : "memory");
  }
}
template <bool kTf32, bool kPair>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc) {
  if constexpr (kTf32 && kPair) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u)
        : "memory");
  } else if constexpr (kTf32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u)
        : "memory");
  } else if constexpr (kPair) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// fp32 reductions into C for split pieces (no return value: fire and forget at the L2)
__device__ __forceinline__ void red_add_v4(float4* dst, const float4& v) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add(float* dst, float v) {
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(dst), "f"(v) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (tcgen05): start address
// >> 4 in [0,14), LBO = 1 in [16,30) (ignored for swizzled K-major), SBO = 1024
// B >> 4 in [32,46) (8 rows x 128 B per swizzle atom), version 1 in [46,48),
// layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}


__device__ __forceinline__ unsigned long long sm_clock() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_put(const SpmmParams& p, bool on, int zone, uint32_t rank,
                                          uint32_t idx, unsigned long long t0,
                                          unsigned long long t1) {
  if (on && idx < static_cast<uint32_t>(p.trace_cap)) {
    unsigned long long* r =
        p.trace + ((static_cast<size_t>(zone) * 2 + rank) * p.trace_cap + idx) * 2;
    r[0] = t0;
    r[1] = t1;
  }
}


__device__ __forceinline__ Item load_item(const SpmmParams& p, int it) {
  const int4 raw = __ldg(reinterpret_cast<const int4*>(p.items) + p.cta_items[it]);
  Item item;
  item.srow = raw.x; item.j0 = raw.y; item.chunk_off = raw.z; item.count = static_cast<uint32_t>(raw.w);
  return item;
}

// ------------------------------------------------------------------ kernel
// kPair = false: one CTA owns a (super-row, 128-column tile) item; MMAs are cta_group::1, M = 128.
// kPair = true : the two CTAs of a cluster (one TPC) own a (super-row, 256-column tile) item.
//   Each CTA stages ITS 128 columns of the B panel and HALF of the rows of every A image run;
//   the leader CTA (cluster rank 0) issues cta_group::2 MMAs with M = 256 that read both
//   CTAs' shared memory, and each CTA's TMEM receives the accumulators of its own 128 columns.
//   Per unit of tensor work an SM therefore pulls half as many A bytes out of L2.
//   Cross-CTA signalling: the peer forwards "my stage is full" to the leader with a remote
//   mbarrier arrive; stage release and accumulator-ready are tcgen05.commit multicasts;
//   "accumulator drained" of the peer's epilogue warps is a remote arrive on the leader.
// kSlots = true : every pipeline stage owns a FIXED slot for its A images (a_slot_bytes, the largest
//   chunk of the handle) next to its B panel.  No ring bookkeeping, so the producer's per-chunk
//   instruction chain is a third as long, and a second producer warp (warp 6) can take every other
//   chunk without sharing any state with the first.  Chosen when at least 3 such stages fit.
// kSlots = false: the A images of the stages in flight share one byte ring (more stages in flight
//   when chunk sizes vary a lot); single producer.
// kWide = true (fixed slots only): a work item covers p.tiles (2 or 4) column tiles, every stage carries that many
//   panels of B.  A separate instantiation because the loops over the tiles sit on the two tightest
//   single-thread chains of the kernel (copy issue, MMA issue): with a run-time count of 1 they cost the
//   one-tile schedules 3 % (bf16) to 9 % (tf32) -- measured, profiles/r2_wide_items.md.
template <bool kTf32, bool kPair, bool kSlots, bool kWide>
__global__ void __launch_bounds__(kSpmmThreads, 1)
spmm_vbr_sm100(const __grid_constant__ CUtensorMap tmap_b, const SpmmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles.
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int P = p.panel_stages;
  const int T = (kSlots && kWide) ? p.tiles : 1;      // B panels per stage (column tiles per work item)
  const uint32_t stage_panels = static_cast<uint32_t>(T) * kPanelBytes;
  const uint32_t panels = base;
  const uint32_t a_ring = base + P * stage_panels;
  uint8_t* ctrl = smem + P * stage_panels + p.a_ring_bytes;
  const uint32_t ctrl_u = a_ring + p.a_ring_bytes;
  // ctrl layout: full[8] | empty[8] | acc_full[2] | acc_empty[2] | tmem_ptr | starts[8] |
  //              run tables[8] (kTableBytes each, filled by the copy engine)
  const uint32_t bar_full = ctrl_u;
  const uint32_t bar_empty = ctrl_u + 64;
  const uint32_t bar_acc_full = ctrl_u + 128;
  const uint32_t bar_acc_empty = ctrl_u + 144;
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(ctrl + 160);
  uint32_t* starts = reinterpret_cast<uint32_t*>(ctrl + 192);
  const uint32_t tables_u = ctrl_u + 256;
  const uint8_t* tables_s = ctrl + 256;
  static_assert(256 + kMaxPanelStages * kTableBytes <= kSmemStageOff, "control block too small");
  static_assert(kSmemStageOff + 4 * 2048 <= kSmemCtrlBytes, "no room for the epilogue staging tiles");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;   // 0 = leader
  const int worker = kPair ? (blockIdx.x >> 1) : blockIdx.x;
  constexpr int kShare = kPair ? 1 : 0;                   // per-CTA bytes = chunk bytes >> kShare

  if (threadIdx.x == 0) {
    for (int s = 0; s < P; ++s) {
      // leader of a pair: own producer + the peer's "my half has landed" relay
      mbar_init(bar_full + 8 * s, (kPair && rank == 0) ? 2 : 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_acc_full + 8 * s, 1);
      mbar_init(bar_acc_empty + 8 * s, kPair ? 8 : 4);  // one arrival per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(const_cast<uint32_t*>(tmem_ptr_s))),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(const_cast<uint32_t*>(tmem_ptr_s))),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  // Programmatic dependent launch (back-to-back multiplies of one handle): the next launch's CTAs may take
  // an SM as soon as this grid's CTA leaves it, and run their copies and MMAs (reads of A, B and the
  // schedule only) while the slowest CTAs of this grid finish; their epilogue warps wait for this grid
  // to complete (griddepcontrol.wait below) before the first write to C.  No-ops in an ordinary launch.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int it_begin = p.cta_ptr[worker];
  const int it_end = p.cta_ptr[worker + 1];
  const bool tr = p.trace != nullptr && worker == p.trace_worker;

  if (kSlots && (warp == 0 || warp == 6)) {
    // ===================== TMA producers, fixed slots (warps 0 and 6, converged) =====================
    // Producer `my` of `step` takes the uses u = my, my + step, ... of the CTA's chunk sequence
    // (a use = one chunk of one item, counted across items); use u lives in stage u % P.  Chunk
    // records are fetched 32 of MY chunks at a time by the whole warp and broadcast by shuffle.
    const uint32_t step = static_cast<uint32_t>(p.producers);
    const uint32_t my = warp == 0 ? 0u : 1u;
    if (my < step) {
      const uint32_t slot_bytes = static_cast<uint32_t>(p.a_slot_bytes);
      uint32_t u = my, slot = my % static_cast<uint32_t>(P), phase = (my / static_cast<uint32_t>(P)) & 1u;
      uint32_t u_base = 0;   // uses of the items before this one
      for (int it = it_begin; it < it_end; ++it) {
        const Item item = load_item(p, it);
        const SuperRow sr = p.srows[item.srow];
        const int j0 = item.j0 + static_cast<int>(rank) * kTileJ;
        const int4* recs = reinterpret_cast<const int4*>(p.chunks + sr.chunk_begin + item.chunk_off);
        const int chunk_count = static_cast<int>(item.count & kItemCountMask);
        const int first = static_cast<int>((my + step - (u_base % step)) % step);   // my first chunk of this item
        const int mine_n = chunk_count > first ? (chunk_count - first + static_cast<int>(step) - 1) / static_cast<int>(step) : 0;
        int4 nxt = make_int4(0, 0, 0, 0);
        int2 nxt_t = make_int2(0, 0);
        if (lane < mine_n) {
          const int c = first + static_cast<int>(step) * lane;
          nxt = __ldg(recs + 2 * c);
          const int4 hi = __ldg(recs + 2 * c + 1);
          nxt_t = make_int2(hi.y, hi.z);
        }
        for (int m0 = 0; m0 < mine_n; m0 += 32) {
          const int4 cur = nxt;
          const int2 cur_t = nxt_t;
          if (m0 + 32 + lane < mine_n) {
            const int c = first + static_cast<int>(step) * (m0 + 32 + lane);
            nxt = __ldg(recs + 2 * c);
            const int4 hi = __ldg(recs + 2 * c + 1);
            nxt_t = make_int2(hi.y, hi.z);
          }
          const int batch = min(32, mine_n - m0);
          for (int i = 0; i < batch; ++i) {
            const int ch_k0 = __shfl_sync(0xFFFFFFFFu, cur.x, i);
            const uint32_t ch_off16 = static_cast<uint32_t>(__shfl_sync(0xFFFFFFFFu, cur.z, i));
            const uint32_t bytes = static_cast<uint32_t>(__shfl_sync(0xFFFFFFFFu, cur.w, i)) >> kShare;
            const uint32_t tb_bytes = static_cast<uint32_t>(__shfl_sync(0xFFFFFFFFu, cur_t.x, i));
            const uint32_t tb_off16 = static_cast<uint32_t>(__shfl_sync(0xFFFFFFFFu, cur_t.y, i));
            const unsigned long long tp0 = tr ? sm_clock() : 0ull;
            if (u >= static_cast<uint32_t>(P)) mbar_wait(bar_empty + 8 * slot, phase ^ 1u, 1);
            if (elect_one()) {
              const uint32_t full = bar_full + 8 * slot;
              const uint8_t* src = p.a_packed + static_cast<size_t>(ch_off16) * 16 + static_cast<size_t>(rank) * bytes;
              const uint32_t dst = a_ring + slot * slot_bytes;
              mbar_arrive_expect_tx(full, stage_panels + bytes + tb_bytes);
              for (int t = 0; t < T; ++t)   // tile t of the item: columns j0 + t * (tile width); beyond n the box is zero-filled
                tma_load_2d(panels + slot * stage_panels + t * kPanelBytes, &tmap_b, ch_k0,
                            j0 + t * (kPair ? 2 * kTileJ : kTileJ), full);
              for (uint32_t done = 0; done < bytes; done += 32768u) {
                const uint32_t piece = min(32768u, bytes - done);
                bulk_load(dst + done, src + done, piece, full);
              }
              bulk_load(tables_u + slot * kTableBytes, p.tables + static_cast<size_t>(tb_off16) * 16, tb_bytes, full);
              if (tr) trace_put(p, true, 0, rank, u, tp0, sm_clock());
            }
            __syncwarp();
            u += step;
            slot += step;
            if (slot >= static_cast<uint32_t>(P)) { slot -= static_cast<uint32_t>(P); phase ^= 1u; }
          }
        }
        u_base += static_cast<uint32_t>(chunk_count);
      }
    }
  } else if (!kSlots && warp == 0) {
    // ===================== TMA producer, byte ring (warp 0 of every CTA, converged) =====================
    // Chunk records are fetched 32 at a time by the whole warp (one coalesced request per
    // batch, the next batch in flight while this one is issued) and broadcast by shuffle: a
    // dependent global load per chunk would cap the issue rate at one chunk per L2 round trip.
    // Per chunk the warp only waits for a stage, places the images in the ring and fires three
    // copies (B panel, A images, run table); everything the MMA warp needs was decoded on the host.
    const uint32_t RB = static_cast<uint32_t>(p.a_ring_bytes);
    uint32_t iss = 0, rel = 0, head = 0;           // uses issued / uses known released / ring head
    uint32_t iss_slot = 0, rel_slot = 0, rel_phase = 0;
    auto release_one = [&](int tag) {              // observe the release of the oldest use
      mbar_wait(bar_empty + 8 * rel_slot, rel_phase, tag);
      ++rel;
      if (++rel_slot == static_cast<uint32_t>(P)) { rel_slot = 0; rel_phase ^= 1u; }
    };
    for (int it = it_begin; it < it_end; ++it) {
      const Item item = load_item(p, it);
      const SuperRow sr = p.srows[item.srow];
      const int j0 = item.j0 + static_cast<int>(rank) * kTileJ;
      const int4* recs = reinterpret_cast<const int4*>(p.chunks + sr.chunk_begin + item.chunk_off);
      const int chunk_count = static_cast<int>(item.count & kItemCountMask);   // chunks of this pass
      int4 nxt = make_int4(0, 0, 0, 0);
      int2 nxt_t = make_int2(0, 0);
      if (lane < chunk_count) {
        nxt = __ldg(recs + 2 * lane);
        const int4 hi = __ldg(recs + 2 * lane + 1);
        nxt_t = make_int2(hi.y, hi.z);             // tbl_bytes, tbl_off16
      }
      for (int c0 = 0; c0 < chunk_count; c0 += 32) {
        const int4 cur = nxt;
        const int2 cur_t = nxt_t;
        if (c0 + 32 + lane < chunk_count) {
          nxt = __ldg(recs + 2 * (c0 + 32 + lane));
          const int4 hi = __ldg(recs + 2 * (c0 + 32 + lane) + 1);
          nxt_t = make_int2(hi.y, hi.z);
        }
        const int batch = min(32, chunk_count - c0);
        for (int i = 0; i < batch; ++i) {
          const int ch_k0 = __shfl_sync(0xFFFFFFFFu, cur.x, i);
          const uint32_t ch_off16 = static_cast<uint32_t>(__shfl_sync(0xFFFFFFFFu, cur.z, i));
          const uint32_t ch_bytes = static_cast<uint32_t>(__shfl_sync(0xFFFFFFFFu, cur.w, i));
          const uint32_t tb_bytes = static_cast<uint32_t>(__shfl_sync(0xFFFFFFFFu, cur_t.x, i));
          const uint32_t tb_off16 = static_cast<uint32_t>(__shfl_sync(0xFFFFFFFFu, cur_t.y, i));
          const unsigned long long tp0 = tr ? sm_clock() : 0ull;
          // stage slot: the use that last occupied it must have been released
          while (iss - rel >= static_cast<uint32_t>(P)) release_one(1);
          // contiguous space in the A ring (FIFO release order).  The offsets depend only on
          // the sequence of sizes, so both CTAs of a pair place every chunk at the same offset.
          const uint32_t bytes = ch_bytes >> kShare;
          uint32_t off;
          for (;;) {
            if (rel == iss) { off = 0; break; }
            const uint32_t tail = starts[rel_slot];
            if (head > tail) {
              if (head + bytes <= RB) { off = head; break; }
              if (bytes <= tail) { off = 0; break; }
            } else if (head + bytes <= tail) {
              off = head;
              break;
            }
            release_one(2);
          }
          const uint32_t s = iss_slot;
          const uint32_t full = bar_full + 8 * s;
          const uint8_t* src = p.a_packed + static_cast<size_t>(ch_off16) * 16 +
                               static_cast<size_t>(rank) * bytes;
          __syncwarp();   // every lane has read starts[] before it is overwritten
          if (elect_one()) {
            starts[s] = off;
            mbar_arrive_expect_tx(full, kPanelBytes + bytes + tb_bytes);
            tma_load_2d(panels + s * kPanelBytes, &tmap_b, ch_k0, j0, full);
            for (uint32_t done = 0; done < bytes; done += 32768u) {
              const uint32_t piece = min(32768u, bytes - done);
              bulk_load(a_ring + off + done, src + done, piece, full);
            }
            bulk_load(tables_u + s * kTableBytes, p.tables + static_cast<size_t>(tb_off16) * 16,
                      tb_bytes, full);
            if (tr) trace_put(p, true, 0, rank, iss, tp0, sm_clock());
          }
          __syncwarp();
          head = off + bytes;
          ++iss;
          if (++iss_slot == static_cast<uint32_t>(P)) iss_slot = 0;
        }
      }
    }
  } else if (warp == 1) {
    if (kPair && rank != 0) {
      // ===================== peer CTA: forward "stage full" to the leader =====================
      const uint32_t remote = map_to_cta(bar_full, 0);
      uint32_t slot = 0, phase = 0;
      for (int it = it_begin; it < it_end; ++it) {
        const int chunk_count = static_cast<int>(load_item(p, it).count & kItemCountMask);
        for (int c = 0; c < chunk_count; ++c) {
          mbar_wait(bar_full + 8 * slot, phase, 6);
          if (lane == 0) mbar_arrive_remote_relaxed(remote + 8 * slot);
          __syncwarp();
          if (++slot == static_cast<uint32_t>(P)) { slot = 0; phase ^= 1u; }
        }
      }
    } else {
      // ===================== MMA issuer (leader CTA, converged warp) =====================
      uint32_t use = 0, slot = 0, phase = 0;
      uint32_t acc_use[2] = {0, 0};
      int local = 0;
      // chained launches, prefetch-only flavour: the copies of the first stages run ahead, the MMAs start
      // together once the previous grid is done (the workers of a team stay in step on their A images)
      if (p.chain_wait_mma) asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int it = it_begin; it < it_end; ++it, ++local) {
        const Item item = load_item(p, it);
        const int chunk_count = static_cast<int>(item.count & kItemCountMask);
        const int as = (p.acc_stages == 2) ? (local & 1) : 0;
        const unsigned long long ta0 = tr ? sm_clock() : 0ull;
        mbar_wait(bar_acc_empty + 8 * as, acc_use[as] & 1, 3);
        if (tr && lane == 0) trace_put(p, true, 3, 0, static_cast<uint32_t>(local), ta0, sm_clock());
        ++acc_use[as];
        tc_fence_after();
        const uint32_t acc_base = tmem_base + as * p.acc_stage_cols;
        for (int c = 0; c < chunk_count; ++c, ++use) {
          const uint32_t s = slot;
          mbar_wait(bar_full + 8 * s, phase, 4);   // pair: own half AND the peer's relay
          const unsigned long long tm0 = tr ? sm_clock() : 0ull;
          tc_fence_after();
          // One elected lane does the whole chunk: it reads the host-decoded run table the copy
          // engine placed in this stage, fires ksteps MMAs per run and commits.  Staying inside a
          // single elect region keeps every operand on the uniform datapath and avoids a
          // shuffle / re-election per run.
          if (elect_one()) {
            const uint4* tbl = reinterpret_cast<const uint4*>(tables_s + s * kTableBytes);
            const uint4 hdr = tbl[0];            // nruns, ksteps
            const uint4 r01 = tbl[1];            // runs 0 and 1: {idesc, where} x 2
            const uint4 r23 = tbl[2];            // runs 2 and 3 (slot is always kTableBytes long)
            const uint32_t a_base = kSlots ? a_ring + s * static_cast<uint32_t>(p.a_slot_bytes) : a_ring + starts[s];
            const uint64_t pdesc0 = smem_desc(panels + s * stage_panels);
            const int nruns = static_cast<int>(hdr.x);
            const int ksteps = static_cast<int>(hdr.y);
            auto fire = [&](uint32_t idesc, uint32_t where) {
              const uint64_t adesc = smem_desc(a_base + ((where & 0xFFFFu) << 4));
              for (int t = 0; t < T; ++t) {      // the same rows of A against every B panel of the stage
                const uint64_t pdesc = pdesc0 + static_cast<uint64_t>(t) * (kPanelBytes >> 4);
                const uint32_t d_tmem = acc_base + static_cast<uint32_t>(t) * static_cast<uint32_t>(p.acc_stage_cols) + (where >> 16);
                if (ksteps == 4) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)   // +32 bytes along K inside the 128-byte swizzle row
                    tc_mma<kTf32, kPair>(d_tmem, pdesc + 2 * k, adesc + 2 * k, idesc);
                } else {
                  for (int k = 0; k < ksteps; ++k)
                    tc_mma<kTf32, kPair>(d_tmem, pdesc + 2 * k, adesc + 2 * k, idesc);
                }
              }
            };
            fire(r01.x, r01.y);
            if (nruns > 1) fire(r01.z, r01.w);
            if (nruns > 2) fire(r23.x, r23.y);
            if (nruns > 3) fire(r23.z, r23.w);
            for (int r = 4; r < nruns; ++r) {
              const uint2 rr = *reinterpret_cast<const uint2*>(tables_s + s * kTableBytes + 16 + 8 * r);
              fire(rr.x, rr.y);
            }
            tc_commit<kPair>(bar_empty + 8 * s);
          }
          __syncwarp();
          if (tr && lane == 0) {   // rank-0 slot: stage seen full; rank-1 slot: {seen full, MMAs issued}
            trace_put(p, true, 1, 0, use, tm0, tm0);
            trace_put(p, true, 1, 1, use, tm0, sm_clock());
          }
          if (++slot == static_cast<uint32_t>(P)) { slot = 0; phase ^= 1u; }
        }
        if (elect_one()) tc_commit<kPair>(bar_acc_full + 8 * as);
        __syncwarp();
      }
      __syncwarp();
    }
  } else if (warp >= 2 && warp <= 5) {
    // ===================== epilogue (warps 2..5, every CTA) =====================
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t acc_empty_remote = kPair ? map_to_cta(bar_acc_empty, 0) : 0u;
    // Zero every accumulator once; the MMAs always accumulate and the epilogue
    // re-zeroes what it drains.
    for (int c0 = 0; c0 < 512; c0 += 16) tmem_st16_zero(t_lane + c0);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (rank == 0) {
        mbar_arrive(bar_acc_empty);
        if (p.acc_stages == 2) mbar_arrive(bar_acc_empty + 8);
      } else {
        mbar_arrive_remote_relaxed(acc_empty_remote);
        if (p.acc_stages == 2) mbar_arrive_remote_relaxed(acc_empty_remote + 8);
      }
    }
    uint32_t acc_use[2] = {0, 0};
    int local = 0;
    // everything below writes (or, with accumulate, reads) C: the previous grid on the stream must be done
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool c_aligned = (reinterpret_cast<uintptr_t>(p.C) & 15) == 0;
    bool zero_seen = p.n_zero_jobs == 0;
    if (p.n_zero_jobs > 0) {
      // this CTA's share of the tiles that split pieces add into (see SpmmParams::zero_jobs);
      // it runs while the producer and MMA warps work on the first item
      const int et = static_cast<int>(threadIdx.x) - 64;   // 0..127 over the four epilogue warps
      const int kHalves = (kPair ? 2 : 1) * T;   // 128-column pieces of one (wide) tile
      for (int u = blockIdx.x; u < p.n_zero_jobs * kHalves; u += gridDim.x) {
        const ZeroJob job = p.zero_jobs[u / kHalves];
        const int j0 = job.j0 + (u % kHalves) * kTileJ;
        const int jn = min(kTileJ, p.n - j0);
        if (jn <= 0) continue;
        const SuperRow zr = p.srows[job.srow];
        for (int sidx = 0; sidx < zr.seg_count; ++sidx) {
          const Segment sg = p.segs[zr.seg_begin + sidx];
          if (c_aligned && p.c_sr == 1 && (p.c_sj & 3) == 0 && (sg.c_row0 & 3) == 0 && (sg.h & 3) == 0) {
            const int q4 = sg.h >> 2;   // 16-byte pieces per column; lanes run along the rows
            for (int i = et; i < jn * q4; i += 128) {
              const int jj = i / q4, r4 = i - jj * q4;
              reinterpret_cast<float4*>(p.C + static_cast<int64_t>(j0 + jj) * p.c_sj + sg.c_row0)[r4] =
                  make_float4(0.f, 0.f, 0.f, 0.f);
            }
          } else {
            for (int i = et; i < jn * sg.h; i += 128) {
              int r, jj;
              if (p.c_sr == 1) { jj = i / sg.h; r = i - jj * sg.h; } else { r = i / jn; jj = i - r * jn; }
              p.C[static_cast<int64_t>(sg.c_row0 + r) * p.c_sr + static_cast<int64_t>(j0 + jj) * p.c_sj] = 0.0f;
            }
          }
        }
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        __threadfence();
        atomicAdd(p.sync_counter, 1ULL);
      }
    }
    float* stg = reinterpret_cast<float*>(ctrl + kSmemStageOff) + (warp - 2) * 512;   // 32 x 16 floats
    for (int it = it_begin; it < it_end; ++it, ++local) {
      const Item item = load_item(p, it);
      const SuperRow sr = p.srows[item.srow];
      const int as = (p.acc_stages == 2) ? (local & 1) : 0;
      // bounded accumulation chains: passes other than the last fold their result into the master
      // copy of the accumulator (TMEM columns master_col..) instead of writing C
      const bool fold_in = (item.count & kItemNotFirst) != 0;
      const bool to_master = (item.count & kItemNotLast) != 0;
      const bool add_c = p.accumulate != 0;
      const bool red_c = (item.count & kItemAtomic) != 0;   // a split piece: C += partial sums
      if (red_c && !zero_seen) {
        // every CTA must have zeroed its share of the split tiles before the first reduction
        if (lane == 0) {
          const uint64_t t0 = global_timer_ns();
          unsigned long long seen;
          uint32_t spins = 0;
          for (;;) {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(p.sync_counter) : "memory");
            if (seen >= p.sync_target) break;
            if (((++spins) & 0xFF) == 0 && global_timer_ns() - t0 > 5000000000ull) {
              printf("sparta spmm: split tiles were not zeroed in time (cta %d, %llu of %llu)\n", blockIdx.x, seen,
                     p.sync_target);
              __trap();
            }
            __nanosleep(200);
          }
        }
        __syncwarp();
        zero_seen = true;
      }
      mbar_wait(bar_acc_full + 8 * as, acc_use[as] & 1, 5);
      const unsigned long long te0 = (tr && warp == 2 && lane == 0) ? sm_clock() : 0ull;
      ++acc_use[as];
      tc_fence_after();
      // (tile 0 of the item; a wide item drains its other column tiles below, updating these four)
      uint32_t t_acc = t_lane + as * p.acc_stage_cols;
      int j = item.j0 + static_cast<int>(rank) * kTileJ + q * 32 + lane;
      bool jv = j < p.n;
      float* cj = p.C + static_cast<int64_t>(j) * p.c_sj;
      // The segment records of the super-row, one per lane, broadcast by shuffle below: a global
      // load per segment inside the loop put an L2 round trip on the critical path of the drain.
      int4 seg_mine = make_int4(0, 0, 0, 0);
      if (lane < sr.seg_count) seg_mine = __ldg(reinterpret_cast<const int4*>(p.segs + sr.seg_begin) + lane);
      auto segment = [&](int sidx) {
        Segment sg;
        sg.c_row0 = __shfl_sync(0xFFFFFFFFu, seg_mine.x, sidx);
        sg.h = __shfl_sync(0xFFFFFFFFu, seg_mine.y, sidx);
        sg.h_pad = __shfl_sync(0xFFFFFFFFu, seg_mine.z, sidx);
        sg.tmem_col = __shfl_sync(0xFFFFFFFFu, seg_mine.w, sidx);
        return sg;
      };
      // 16 accumulator columns (rows c0.. of segment sg, column j of C per lane) -> C
      auto store16 = [&](const uint32_t (&v)[16], const Segment& sg, int c0) {
        const bool vec_ok = c_aligned && p.c_sr == 1 && (p.c_sj & 3) == 0 && (sg.c_row0 & 3) == 0;
        if (vec_ok && c0 + 16 <= sg.h) {
          // Column-major C: a thread owns one column j, so its 16 values are 64 contiguous bytes
          // but the warp's 32 columns are 32 different lines.  Transposed through a 2 KB per-warp
          // staging tile (XOR-swizzled, conflict-free both ways) 4 lanes cover a column's 64
          // bytes and one store instruction touches 8 lines instead of 32.
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(stg + lane * 16 + 4 * (g ^ ((lane >> 1) & 3))) =
                make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                            __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
          __syncwarp();
          const int gl = lane & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int jj = 8 * i + (lane >> 2);
            float4 o = *reinterpret_cast<const float4*>(stg + jj * 16 + 4 * (gl ^ ((jj >> 1) & 3)));
            const int jcol = j - lane + jj;
            if (jcol < p.n) {
              float4* d4 = reinterpret_cast<float4*>(p.C + static_cast<int64_t>(jcol) * p.c_sj +
                                                     (sg.c_row0 + c0 + 4 * gl));
              if (red_c) {
                red_add_v4(d4, o);
              } else {
                if (add_c) {
                  const float4 old = *d4;
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *d4 = o;
              }
            }
          }
          __syncwarp();
        } else if (jv) {
          float* dst = cj + static_cast<int64_t>(sg.c_row0 + c0) * p.c_sr;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            if (c0 + r < sg.h) {
              float o = __uint_as_float(v[r]);
              float* d = dst + static_cast<int64_t>(r) * p.c_sr;
              if (red_c) {
                red_add(d, o);
              } else {
                if (add_c) o += *d;
                *d = o;
              }
            }
          }
        }
      };
      if (fold_in || to_master) {
        // bounded accumulation chains (the tf32 default): fold through the master accumulators
        for (int sidx = 0; sidx < sr.seg_count; ++sidx) {
          const Segment sg = segment(sidx);
          for (int c0 = 0; c0 < sg.h_pad; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_acc + sg.tmem_col + c0, v);
            if (fold_in) {
              uint32_t m[16];
              tmem_ld16(t_acc + p.master_col + sg.tmem_col + c0, m);
              tmem_wait_ld();
#pragma unroll
              for (int r = 0; r < 16; ++r)
                v[r] = __float_as_uint(__fadd_rn(__uint_as_float(v[r]), __uint_as_float(m[r])));
            } else {
              tmem_wait_ld();
            }
            tmem_st16_zero(t_acc + sg.tmem_col + c0);
            if (to_master) tmem_st16(t_acc + p.master_col + sg.tmem_col + c0, v);
            else store16(v, sg, c0);
          }
        }
      } else {
        // Plain drain, software-pipelined: the accumulator columns of a super-row are contiguous
        // (tmem_col = running sum of h_pad), step s covers columns [16 s, 16 s + 16).  The load of
        // step s+1 is in flight while step s is stored; tcgen05.wait::ld sits right before the
        // first use of a buffer and nothing touches the buffer between its load and that wait.
        const int nsteps = sr.n_cols >> 4;
        for (int t = 0; t < T; ++t) {
        if (t > 0) {   // next column tile of a wide item: its accumulators sit acc_stage_cols further
          t_acc += static_cast<uint32_t>(p.acc_stage_cols);
          j += kPair ? 2 * kTileJ : kTileJ;
          jv = j < p.n;
          cj = p.C + static_cast<int64_t>(j) * p.c_sj;
        }
        int sidx = 0, c0 = 0;
        Segment sg = segment(0);
        auto advance = [&]() {
          c0 += 16;
          if (c0 >= sg.h_pad && sidx + 1 < sr.seg_count) { ++sidx; sg = segment(sidx); c0 = 0; }
        };
        uint32_t va[16], vb[16];
        if (nsteps > 0) tmem_ld16(t_acc, va);
        for (int st = 0; st < nsteps; st += 2) {
          tmem_wait_ld();
          if (st + 1 < nsteps) tmem_ld16(t_acc + 16 * (st + 1), vb);
          tmem_st16_zero(t_acc + 16 * st);
          store16(va, sg, c0);
          advance();
          if (st + 1 < nsteps) {
            tmem_wait_ld();
            if (st + 2 < nsteps) tmem_ld16(t_acc + 16 * (st + 2), va);
            tmem_st16_zero(t_acc + 16 * (st + 1));
            store16(vb, sg, c0);
            advance();
          }
        }
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(bar_acc_empty + 8 * as);
        else mbar_arrive_remote_relaxed(acc_empty_remote + 8 * as);   // TMEM is ordered by tcgen05 fences
        if (tr && warp == 2) trace_put(p, true, 2, rank, static_cast<uint32_t>(local), te0, sm_clock());
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  if constexpr (kPair) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair) {
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(512u)
                   : "memory");
    } else {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(512u)
                   : "memory");
    }
  }
}

// ------------------------------------------------------------------ launch
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

typedef void (*KernelFn)(const CUtensorMap, const SpmmParams);
static KernelFn pick_kernel(bool tf32, bool pair, bool slots, bool wide = false) {
  if (slots && wide)
    return tf32 ? (pair ? spmm_vbr_sm100<true, true, true, true> : spmm_vbr_sm100<true, false, true, true>)
                : (pair ? spmm_vbr_sm100<false, true, true, true> : spmm_vbr_sm100<false, false, true, true>);
  if (slots)
    return tf32 ? (pair ? spmm_vbr_sm100<true, true, true, false> : spmm_vbr_sm100<true, false, true, false>)
                : (pair ? spmm_vbr_sm100<false, true, true, false> : spmm_vbr_sm100<false, false, true, false>);
  return tf32 ? (pair ? spmm_vbr_sm100<true, true, false, false> : spmm_vbr_sm100<true, false, false, false>)
              : (pair ? spmm_vbr_sm100<false, true, false, false> : spmm_vbr_sm100<false, false, false, false>);
}

// CTAs of the persistent grid the device can hold AT THE SAME TIME at this shared-memory
// footprint (one per SM, in clusters of 2 in pair mode).  The in-kernel zeroing of split tiles
// spins on a grid-wide counter, which is only safe when the whole grid is co-resident: the
// handle clamps its grid to this number (abi.cu, create_common).
cudaError_t spmm_max_coresident_ctas(int pair, int kind_tf32, int slots, int smem_bytes, int* out) {
  KernelFn fn = pick_kernel(kind_tf32 != 0, pair != 0, slots != 0);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2, 1, 1);
  cfg.blockDim = dim3(kSpmmThreads, 1, 1);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int clusters = 0;
  e = cudaOccupancyMaxActiveClusters(&clusters, fn, &cfg);
  if (e != cudaSuccess) return e;
  *out = clusters * (pair ? 2 : 1);
  return cudaSuccess;
}

cudaError_t spmm_launch(const SpmmParams& p, const void* b_dev, int64_t k_total, int64_t ldk,
                        int precision, int grid, cudaStream_t stream, const char** err, bool overlap_previous) {
  *err = "";
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    *err = "cuTensorMapEncodeTiled entry point not available";
    return cudaErrorNotSupported;
  }
  const int esize = prec_esize(precision);
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_total), static_cast<cuuint64_t>(p.n)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldk) * esize};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esize), kTileJ};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = precision == PREC_BF16   ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : precision == PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                          : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUresult r = encode(&tmap, dt, 2, const_cast<void*>(b_dev), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed for the B operand";
    return cudaErrorInvalidValue;
  }
  const int tiles = p.a_slot_bytes > 0 ? p.tiles : 1;
  if (tiles != 1 && tiles != 2 && tiles != 4) { *err = "tiles must be 1, 2 or 4"; return cudaErrorInvalidConfiguration; }
  if (tiles > 1 && (p.acc_stages != 1 || p.acc_stage_cols * tiles != 512 || p.master_col != 0)) {
    *err = "wide items need one accumulator stage of 512 / tiles columns";
    return cudaErrorInvalidConfiguration;
  }
  const int smem = spmm_smem_bytes(p.panel_stages, p.a_ring_bytes, tiles);
  if (smem > kSmemMax || p.panel_stages < 2 || p.panel_stages > kMaxPanelStages) {
    *err = "invalid pipeline configuration (shared memory)";
    return cudaErrorInvalidConfiguration;
  }
  if (p.pair && (grid & 1)) {
    *err = "pair mode needs an even grid";
    return cudaErrorInvalidConfiguration;
  }
  if (p.a_slot_bytes > 0 && (p.a_slot_bytes % 1024 || p.a_slot_bytes * p.panel_stages != p.a_ring_bytes)) {
    *err = "invalid slot configuration";
    return cudaErrorInvalidConfiguration;
  }
  if (p.producers < 1 || p.producers > 2 || (p.producers == 2 && p.a_slot_bytes == 0)) {
    *err = "two copy warps need the fixed-slot pipeline";
    return cudaErrorInvalidConfiguration;
  }
  KernelFn fn = pick_kernel(p.kind_tf32 != 0, p.pair != 0, p.a_slot_bytes > 0, tiles > 1);
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { *err = "cudaFuncSetAttribute(smem)"; return e; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
  cfg.blockDim = dim3(p.producers == 2 ? kSpmmThreads : kSpmmThreads - 32, 1, 1);   // warp 6 exists only as a producer
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.pair ? 2 : 1;   // the two CTAs of a pair share one TPC
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (overlap_previous && p.trace == nullptr) {
    // the kernel before this one on the stream is the same handle's previous multiply (see the kernel prologue)
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  e = cudaLaunchKernelEx(&cfg, fn, tmap, p);
  if (e != cudaSuccess) { *err = "spmm kernel launch"; return e; }
  e = cudaGetLastError();
  if (e != cudaSuccess) *err = "spmm kernel launch";
  return e;
}

}  // namespace sparta
