// Host-side tile scheduler: turns the VBR / Blocked-ELL index arrays into the
// work lists the sm_100a kernel walks.  Pure C++ (no CUDA) so it is testable on
// a CPU-only box.
#pragma once
#include <stdint.h>
#include <vector>
#include "sched_types.h"

namespace sparta {

// Format-neutral view of the nonzero blocks of a range of block-rows.
struct BlockRows {
  int64_t w = 0;                  // column-block width (VBR::block_col_size)
  std::vector<int64_t> row0;      // first C row of each block-row (relative to the shard)
  std::vector<int64_t> height;    // rows in the block-row
  std::vector<int64_t> ptr;       // [count+1] ranges into col/src
  std::vector<int64_t> col;       // column-block index jb of each nonzero block
  std::vector<int64_t> src;       // element offset of the block's (0,0) entry in the fp32 source
  std::vector<int64_t> rs;        // per block-row: element stride between rows of a block
  std::vector<int64_t> ks;        // per block-row: element stride between k of a block
  // Optional per-block overrides (empty = the defaults above).  They describe A TRANSPOSED, the
  // operand of the inverted product C = B*A: its block-rows are A's column blocks and its
  // "column blocks" are A's block-rows, whose extents along k are the variable heights.
  std::vector<int64_t> blk_k0;    // first k of the block (default col * w)
  std::vector<int64_t> blk_kw;    // extent of the block along k (default w)
  std::vector<int64_t> blk_rs;    // element stride between rows of the block (default rs[b])
  std::vector<int64_t> blk_ks;    // element stride between k of the block (default ks[b])
  // Optional sub-block lists (empty = every block is one sub-block covering the whole height).
  // They describe FUSED block-rows (fuse_short_block_rows): a block of a fused block-row is the
  // stack of the blocks its member block-rows own at that column block, each with its own source.
  std::vector<int64_t> sub_ptr;   // [blocks + 1] ranges into the sub_* arrays
  std::vector<int64_t> sub_off;   // first row of the sub-block inside the (fused) block-row
  std::vector<int64_t> sub_h;     // its rows
  std::vector<int64_t> sub_src;   // element offset of its (0,0) entry in the fp32 source
  std::vector<int64_t> sub_rs;    // element stride between its rows
  std::vector<int64_t> sub_ks;    // element stride between its k
  int64_t count() const { return static_cast<int64_t>(height.size()); }
};

// Variable-height blockings (-a 3/4 at low tau) produce long runs of block-rows one or two rows
// tall; each would occupy a 16-row MMA segment of its own and walk its own column-block list.
// This fuses runs of CONSECUTIVE block-rows (adjacent rows of C) whose heights add up to at most
// max_rows into one block-row whose column-block list is the union of theirs; rows without a block
// at a column block are zero rows of the image.  The device images never get larger (a union is
// at most the sum) and usually several times smaller.  The host VBR arrays are untouched and
// FLOPs are still counted on the original blocks.  Returns false (and leaves *out alone) when
// there is nothing to fuse or the view already has per-block overrides.
bool fuse_short_block_rows(const BlockRows& in, int max_rows, BlockRows* out);

struct ScheduleOptions {
  int precision = PREC_BF16;
  int seg_rows = 64;       // max rows per segment (multiple of 16, <= 256)
  int acc_cols = 512;      // TMEM columns per accumulator stage: 256 (2 stages) or 512 (1)
  int tiles = 1;           // column tiles of B one work item covers (1, 2 or 4): every stage then carries `tiles`
                           // panels of B for ONE set of A images, and a super-row gets 512 / tiles accumulator columns
  int num_ctas = 148;      // persistent grid size upper bound
  int pair = 1;            // 1: two CTAs (one TPC) share a super-row through tcgen05 cta_group::2
  int sort_rows = 1;       // 1: group block-rows of similar nonzero-block count into super-rows
  int64_t l2_slab_bytes = 160ll << 20;  // B columns kept L2-resident at a time (see build_assignment)
  int max_chain = 0;       // longest accumulation chain in tcgen05.mma instructions (0: default, -1: unlimited)
  int split = 0;           // chunk lists split between workers: 0 when the model says it pays, 1 never, 2 always
};

struct Structure {           // independent of the number of B columns
  std::vector<Segment>  segs;
  std::vector<SuperRow> srows;
  std::vector<Chunk>    chunks;
  std::vector<std::vector<PackJob>> job_parts;   // the pack jobs, in order, as the scheduler's threads built them
  int64_t n_jobs = 0;
  std::vector<PackJob>  jobs;        // all of them in one array: filled by merge_jobs() on request only
  void merge_jobs() {
    if (!jobs.empty() || n_jobs == 0) return;
    jobs.reserve(static_cast<size_t>(n_jobs));
    for (const auto& part : job_parts) jobs.insert(jobs.end(), part.begin(), part.end());
  }
  std::vector<uint32_t> tables;      // run tables of all chunks, back to back (see sched_types.h)
  std::vector<double>   srow_cost;   // modelled SM cycles per column tile (fixed part + chunks)
  std::vector<float>    chunk_cost;  // modelled SM cycles of every chunk
  std::vector<double>   srow_fixed;  // per super-row: pipeline fill + epilogue part of srow_cost
  int64_t chain = -1;                // accumulation-chain bound the passes were cut for (-1: none)
  std::vector<int32_t>  pass_ptr;    // [srows + 1] ranges into pass_off
  std::vector<int32_t>  pass_off;    // per pass: first chunk (relative to the super-row), see Item
  int master_col = 0;                // > 0: TMEM column of the master accumulators (bounded chains)
  int acc_cols = 512;                // accumulator columns the super-rows were packed for
  int64_t max_chain_seen = 0;        // longest accumulation chain (MMAs) of any pass
  uint64_t a_bytes = 0;              // bytes of packed A images
  int64_t  nztot = 0;                // sum over blocks of h*w (the reference's VBR::nztot)
  int64_t  n_blocks = 0;
  int64_t  rows = 0;                 // C rows covered by the shard
  uint32_t max_chunk_bytes = 0;      // largest per-CTA chunk (what must fit in the smem ring)
  int pair = 0;
  int tiles = 1;                     // column tiles per work item the super-rows were packed for (ScheduleOptions::tiles)
  bool sparse_images = false;        // jobs write only their own rows: the image buffer must be zeroed first
};

// Decomposition of a chunk's member mask into MMA runs.  A run is a maximal group of
// consecutive PRESENT members that does not cross a break (SuperRow::break_mask cuts the members
// into fixed groups of at most 256 accumulator columns, the widest N of one tcgen05.mma).  The
// kernel issues one MMA per K step per run; in pair mode the run's rows are split half/half
// between the two CTAs.  The rule is presence-local on purpose: the kernel's producer warp
// derives the runs of a chunk with one ballot-style expression per lane (spmm_kernel.cu).
// cols[m] = first accumulator column of member m, cols[count] = total.  fn(m_begin, m_end, N).
inline uint32_t run_starts(uint32_t mask, uint32_t break_mask) {
  return mask & (~(mask << 1) | break_mask);
}
template <class F>
inline void for_each_run(uint32_t mask, uint32_t break_mask, const int32_t* cols, F fn) {
  uint32_t starts = run_starts(mask, break_mask);
  while (starts) {
    const int m0 = __builtin_ctz(starts);
    starts &= starts - 1;
    int m1 = m0 + 1;
    while (m1 < 32 && ((mask >> m1) & 1) && !((break_mask >> m1) & 1)) ++m1;
    fn(m0, m1, cols[m1] - cols[m0]);
  }
}

struct Assignment {          // depends on n (columns of B) and the grid
  std::vector<Item>    items;
  std::vector<int32_t> cta_ptr;     // per worker (a CTA, or a CTA pair in pair mode)
  std::vector<int32_t> cta_items;
  int grid = 0;                     // CTAs to launch (2 x workers in pair mode)
  int workers = 0;
  int team = 1;                     // workers walking the same super-row sequence side by side
  int group_tiles = 1;              // column tiles per L2-resident group
  std::vector<ZeroJob> zero_jobs;   // C tiles written by split pieces (zeroed before every launch)
  int split_pieces = 0;             // items that carry kItemAtomic on their last pass
  double max_cta_cost = 0, mean_cta_cost = 0;
};

// Returns "" on success, otherwise a static error string.
const char* build_structure(const BlockRows& br, const ScheduleOptions& opt, Structure* out);
const char* build_assignment(const Structure& st, const ScheduleOptions& opt, int64_t n,
                             int64_t k_total, Assignment* out);

// Contiguous block-row ranges balanced on nonzero-block area (sum h*w), the
// partition SURVEY 8(e) prescribes for multi-GPU sharding.  cuts has parts+1 entries.
void partition_block_rows(int64_t block_rows, const int64_t* row_part, const int64_t* nzcount,
                          int parts, int64_t* cuts);

}  // namespace sparta
