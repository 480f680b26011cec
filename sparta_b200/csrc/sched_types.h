// Tile-schedule records shared by the host scheduler (schedule.cpp) and the
// sm_100a SpMM kernel (spmm_kernel.cu).  All plain 32-bit PODs so they can be
// memcpy'd to the device unchanged.
//
// Vocabulary (follows the reference's VBR terms, include/matrices.h:93-122):
//   block-row  : one entry of VBR::row_part (variable height h)
//   segment    : <= seg_rows consecutive rows of one block-row (MMA N operand)
//   super-row  : a set of segments whose fp32 accumulators share one TMEM
//                allocation; the unit a CTA owns for one column tile of B
//   chunk      : one (column block, 128-byte K slab) of a super-row's merged
//                column-block list; the unit of the smem pipeline
#pragma once
#include <stdint.h>

namespace sparta {

struct Segment {
  int32_t c_row0;     // first row of C (blocked row order) this segment writes
  int32_t h;          // true number of rows
  int32_t h_pad;      // h rounded up to 16 (MMA N granularity at M=128)
  int32_t tmem_col;   // first accumulator column inside the super-row
};

struct SuperRow {
  int32_t seg_begin;    // into segs[]
  int32_t seg_count;    // <= 32 (chunk mask width)
  int32_t chunk_begin;  // into chunks[]
  int32_t chunk_count;
  int32_t n_cols;       // sum of h_pad (<= tmem columns per accumulator stage)
  uint32_t break_mask;  // bit m set: an MMA run may not continue from member m-1 into member m
                        // (members are cut into fixed groups of <= 256 accumulator columns)
  int32_t pad_[2];
};

struct Chunk {
  int32_t  k0;        // first B row (k index) of this K slab
  uint32_t mask;      // bit m set <=> member segment m has a block here
  uint32_t a_off16;   // offset of the packed A images of this chunk, 16-byte units
  uint32_t a_bytes;   // total bytes of those images (sum h_pad*128 over set bits)
  int32_t  ksteps;    // MMA K steps to issue (1..4), covers the block's true width
  uint32_t tbl_bytes; // bytes of this chunk's run table (16 + 8 per run, rounded up to 16)
  uint32_t tbl_off16; // its offset in the table stream, 16-byte units
  int32_t  pad_;
};

// Run table of a chunk: what the MMA-issuing warp executes, decoded on the host and streamed
// into shared memory by the copy engine next to the A images.  Tables are stored back to back:
//   word 0 nruns, word 1 ksteps, words 2-3 zero, then per run
//   .x = tcgen05 instruction descriptor (M, N, formats)
//   .y = accumulator column << 16 | (byte offset of the run's rows inside the CTA's images >> 4)
constexpr int kTableWords = 4 + 2 * 32;
constexpr int kTableBytes = kTableWords * 4;   // 272: the largest table (32 runs)

// A work item = one PASS of a (super-row, column tile): a contiguous range of the super-row's
// chunks accumulated into the working accumulator.  Normally a super-row is one pass.  When the
// accumulation chain is bounded (ScheduleOptions::max_chain, the tf32 default) a long super-row is
// cut into several passes executed back to back by the same worker: the epilogue folds every
// pass into a MASTER copy of the accumulator kept in the other half of TMEM with round-to-nearest
// fp32 adds, and only the last pass writes C.  This bounds the error of the tensor core's
// truncating fp32 accumulation, which otherwise grows linearly with the number of MMAs.
//
// A (super-row, column tile) may also be SPLIT along its chunk list between several workers
// (schedule.cpp, build_assignment): when a shard holds too few super-rows to fill the grid -- one
// rank of an 8-GPU run owns ~10 super-rows x 8 column tiles for 74 CTA pairs -- the chunk lists are
// laid end to end and cut into equal-cost pieces.  A piece that does not cover the whole list adds
// its partial sums to C with fp32 reductions (kItemAtomic); the tiles those pieces write are
// zeroed by a small kernel before the launch (Assignment::zero_jobs).
constexpr uint32_t kItemNotFirst = 1u << 31;   // fold the master copy into this pass's result
constexpr uint32_t kItemNotLast  = 1u << 30;   // store the result to the master copy, not to C
constexpr uint32_t kItemAtomic   = 1u << 29;   // C += result with red.global.add (a split piece)
constexpr uint32_t kItemCountMask = (1u << 29) - 1;
struct Item {
  int32_t srow;       // super-row id
  int32_t j0;         // first column of B/C of this column tile
  int32_t chunk_off;  // first chunk of the pass, relative to SuperRow::chunk_begin
  uint32_t count;     // chunks in the pass | kItemNotFirst | kItemNotLast
};

// A C tile that split pieces accumulate into: the rows of super-row `srow`, columns [j0, j0+tile).
struct ZeroJob {
  int32_t srow;
  int32_t j0;
};

// One packed A image = h_pad rows x 128 bytes in the K-major SWIZZLE_128B
// canonical layout, i.e. the exact bytes the MMA reads from shared memory.
struct PackJob {
  int64_t src_base;   // element offset into the fp32 source (VBR mab / ELL values)
  int64_t src_rs;     // element stride between consecutive rows of the block
  int64_t src_ks;     // element stride between consecutive k of the block
  int32_t h;          // valid rows
  int32_t h_pad;
  int32_t k_lo;       // block-local k of image column 0 (negative when the slab starts
                      // before the block: TMA needs a 16-byte aligned k coordinate)
  int32_t k_w;        // block width w: image column kk holds k = k_lo + kk iff 0 <= k < w
  uint32_t dst_off16; // destination offset, 16-byte units
  int32_t r_base;     // row of the image (dst_off16 = its row 0) that source row 0 goes to
  int32_t pad_[2];
};

enum Precision : int32_t { PREC_BF16 = 0, PREC_FP16 = 1, PREC_TF32 = 2 };

static inline int prec_esize(int p) { return p == PREC_TF32 ? 4 : 2; }

constexpr int kTileJ      = 128;     // MMA M = columns of B per tile
constexpr int kPanelBytes = 16384;   // 128 rows x 128 bytes
constexpr int kMaxMembers = 32;

}  // namespace sparta
