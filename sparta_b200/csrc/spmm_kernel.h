// Launch interface of the sm_100a block-sparse x dense kernel (spmm_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sched_types.h"

namespace sparta {

struct SpmmParams {
  const Item*     items;
  const int32_t*  cta_ptr;      // [grid + 1] ranges into cta_items
  const int32_t*  cta_items;    // item ids grouped per CTA (host LPT assignment)
  const SuperRow* srows;
  const Segment*  segs;
  const Chunk*    chunks;
  const uint8_t*  a_packed;     // packed A images (see PackJob)
  const uint8_t*  tables;       // run tables of all chunks (sched_types.h)
  float*          C;
  int64_t         c_sr;         // element stride of C between rows
  int64_t         c_sj;         // element stride of C between columns
  int32_t         n;            // columns of B and C
  int32_t         accumulate;   // 1: C += A*B (reference beta = 1), 0: C = A*B
  int32_t         kind_tf32;    // 0: kind::f16 (bf16/fp16), 1: kind::tf32
  int32_t         panel_stages; // pipeline depth (<= 8)
  int32_t         a_ring_bytes; // bytes of the A-image ring (multiple of 1024)
  int32_t         acc_stages;   // 1 or 2 TMEM accumulator stages
  int32_t         acc_stage_cols; // 512 / acc_stages
  int32_t         master_col;   // > 0: TMEM column offset of the master accumulators (bounded chains)
  int32_t         pair;         // 1: CTA pairs (cluster of 2, tcgen05 cta_group::2); cta_ptr is per pair
  int32_t         a_slot_bytes; // > 0: fixed-slot pipeline, every stage owns this many bytes for its A images
                                //      (a_ring_bytes = panel_stages * a_slot_bytes); 0: byte ring
  int32_t         producers;    // copy-issuing warps per CTA: 1, or 2 (fixed slots only; alternate chunks)
  int32_t         chain_wait_mma; // launched with programmatic stream serialisation: 1 = the MMA warp waits for the
                                // previous grid before its first MMA (only the first stages' copies overlap that
                                // grid's tail), 0 = only the epilogue warps wait (a whole item's MMAs may overlap)
  int32_t         tiles;        // column tiles of B per work item (fixed slots only): every stage holds `tiles` B panels
                                // for one set of A images; tile t accumulates in TMEM columns [t * 512 / tiles, ...)
  // Split pieces (kItemAtomic) add into C tiles that must start from zero.  Every CTA zeroes its
  // share of those tiles in its epilogue warps while its first item is still in the tensor pipe,
  // then bumps *sync_counter; a warp about to issue its first reduction waits until the counter
  // has reached sync_target (= all CTAs of all launches so far).  All CTAs of the grid are
  // co-resident (one per SM), so the wait cannot deadlock.
  const ZeroJob*  zero_jobs;
  int32_t         n_zero_jobs;
  unsigned long long* sync_counter;
  unsigned long long  sync_target;
  // optional timeline of one worker (sparta_run_traced): 4 zones x 2 ranks x trace_cap records of
  // {t0, t1} SM clocks; zone 0 producer per chunk, 1 MMA per chunk, 2 epilogue per item,
  // 3 MMA accumulator wait per item
  unsigned long long* trace;
  int32_t         trace_worker;
  int32_t         trace_cap;
};

constexpr int kSpmmThreads   = 224;   // warp0 TMA, warp1 MMA, warps2-5 epilogue, warp6 second TMA producer
constexpr int kMaxPanelStages = 8;
constexpr int kSmemStageOff  = 3072;  // offset of the epilogue warps' staging tiles in the control block
constexpr int kSmemCtrlBytes = 3072 + 4 * 2048;  // barriers + per-stage run tables + 4 staging tiles
constexpr int kSmemMax       = 232448; // 227 KB opt-in limit per CTA

// Bytes of dynamic shared memory for a configuration (includes 1 KB slack used
// to align the base to 1024 for SWIZZLE_128B).
static inline int spmm_smem_bytes(int panel_stages, int a_ring_bytes, int tiles = 1) {
  return 1024 + panel_stages * tiles * kPanelBytes + a_ring_bytes + kSmemCtrlBytes;
}

// B operand as the kernel reads it: [n][ldk] elements (k contiguous), converted
// to the compute precision.  Returns cudaSuccess or the failing status; *err
// gets a static description on failure.
cudaError_t spmm_launch(const SpmmParams& p, const void* b_dev, int64_t k_total,
                        int64_t ldk, int precision, int grid, cudaStream_t stream,
                        const char** err, bool overlap_previous = false);
// overlap_previous: launch with programmatic stream serialisation -- allowed only when the previous kernel on
// the stream reads and writes nothing this launch reads before its epilogue (the same handle's previous
// multiply: A, B and the schedule are read-only between set_B calls).

// CTAs of the persistent grid that can be resident at the same time (see spmm_kernel.cu).
cudaError_t spmm_max_coresident_ctas(int pair, int kind_tf32, int slots, int smem_bytes, int* out);

}  // namespace sparta
