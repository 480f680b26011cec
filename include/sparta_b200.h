/* sparta_b200 -- C ABI of the B200 (sm_100a) block-sparse x dense multiply.
 *
 * Drop-in boundary for the SpMM hot path of HicrestLaboratory/SPARTA.  Each
 * entry point states the reference interface it replaces (paths relative to the
 * reference checkout).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions shared with the reference (include/definitions.h:4-6):
 *   intT  = int64_t ("long"),  DataT = DataT_C = float.
 *   C = A*B with alpha = beta = 1 on a caller-zeroed C (test/cuda/cuda_multiply.cpp:134).
 *   The reference never uploads C (cuda_utilities.cpp:95-105), so "accumulate = 0"
 *   (C := A*B) is the drop-in behaviour; accumulate = 1 gives a strict C += A*B.
 *   C rows come back in BLOCKED row order (row_part order), exactly like
 *   VBR::multiply (src/general/vbr.cpp:355).
 *
 * Every function returns 0 on success and a nonzero code otherwise;
 * sparta_last_error() returns a description for the calling thread.  There is
 * no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SPARTA_B200_H
#define SPARTA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPARTA_ABI_VERSION 1

/* operand precision of the tensor-core path; accumulation is always fp32 */
enum { SPARTA_BF16 = 0, SPARTA_FP16 = 1, SPARTA_TF32 = 2 };
/* dense operand layouts (0 = the handle type's default) */
enum { SPARTA_LAYOUT_DEFAULT = 0, SPARTA_COL_MAJOR = 1, SPARTA_ROW_MAJOR = 2 };
/* status codes */
enum {
  SPARTA_OK = 0,
  SPARTA_ERR_INVALID = 1,   /* bad argument */
  SPARTA_ERR_CUDA = 2,      /* CUDA runtime / driver failure (message has the cause) */
  SPARTA_ERR_NO_DEVICE = 3, /* no sm_100 device visible */
  SPARTA_ERR_STATE = 4      /* call order violated (e.g. run before set_B) */
};

typedef struct sparta_handle sparta_handle;

/* Tunables.  Zero-initialise and set struct_size = sizeof(sparta_options);
 * every field left 0 takes its default. */
typedef struct sparta_options {
  int32_t struct_size;
  int32_t precision;     /* SPARTA_BF16 (default) | SPARTA_FP16 | SPARTA_TF32 */
  int32_t device;        /* 1 + CUDA device ordinal; 0 = the current device */
  int32_t b_layout;      /* default: COL_MAJOR for VBR handles, ROW_MAJOR for BELLPACK/CSR */
  int32_t c_layout;      /* same default as b_layout */
  int32_t accumulate;    /* 0: C := A*B (default) ; 1: C += A*B */
  int32_t seg_rows;      /* max rows per MMA segment, multiple of 16 <= 256 (default 64) */
  int32_t acc_cols;      /* TMEM columns per accumulator stage: 512 (default, one stage) or 256 (two) */
  int32_t panel_stages;  /* smem pipeline depth, 2..8 (default 5) */
  int32_t num_ctas;      /* persistent grid size (default: SM count) */
  int64_t block_row_begin; /* shard: first block-row (default 0) */
  int64_t block_row_end;   /* shard: one past the last block-row.  With explicit_range = 0 a value <= 0 means
                              "to the end" (so a zeroed struct selects everything); set explicit_range = 1 to
                              have [begin, end) taken literally, including an EMPTY shard (begin == end) */
  int32_t cta_pair;      /* 0/2: CTA pairs, tcgen05 cta_group::2, 256-column tiles (default); 1: single CTAs */
  int32_t row_order;     /* 0/2: super-rows group block-rows of similar block count (default); 1: input order */
  int32_t l2_slab_mb;    /* B columns walked per pass over A, in MiB of B (default 160) */
  int32_t max_chain;     /* longest run of tcgen05.mma accumulations into one TMEM accumulator before the
                            partial sum is drained and added to C in fp32 by the epilogue; 0 = the
                            precision's default (tf32: 256 per accumulator, bf16/fp16: unlimited), -1 = unlimited */
  int32_t split_k;       /* few super-rows per worker (small shards): cut the block-rows' column-block
                            lists into equal-cost pieces, one per worker, partial sums added to C with
                            fp32 reductions.  0: when the cost model says it pays (default), 1: never,
                            2: always */
  int32_t fuse_rows;     /* runs of consecutive block-rows whose heights add up to <= 16 share one 16-row MMA
                            segment with the union of their column-block lists (variable-height blockings
                            produce thousands of block-rows one or two rows tall).  0: on (default), 1: off */
  int32_t explicit_range; /* 1: block_row_begin / block_row_end are literal (an empty range is an empty shard) */
  int32_t pipeline;      /* shared-memory pipeline of the SpMM kernel.  0: fixed slots when at least 3 stages of
                            (B panel + the handle's largest chunk of A images) fit, else the byte ring (default);
                            1: byte ring (A images of the stages in flight share one ring; one copy warp);
                            2: fixed slots */
  int32_t copy_warps;    /* fixed slots only: warps per CTA that issue the TMA / bulk copies, taking alternate
                            chunks.  0/2: two (default; one thread cannot start more than one pipeline stage per
                            ~600 cycles, profiles/r2_copy_issue_microbench.txt), 1: one */
  int32_t gather_max_height; /* VBR handles: block-rows of at most this many rows do not go through the tensor
                            cores (an MMA needs 8-16 rows of N; a block-row of height 1 would be 94 % padding and
                            fetch a 16 KB panel of B to use one row of it) but through the gather kernel of the
                            family (csr_kernel.cu) on the NONZEROS of their blocks -- same numbers, operands rounded
                            to the handle's precision, fp32 accumulation.  0: default (7), -1: off, k > 0: heights <= k.
                            Variable-height blockings (-a 3 / -a 4) leave most block-rows one row tall. */
  int32_t gather_passes;  /* launches of the gather kernel per multiply, each over one range of A's columns (the rows
                            of B a launch reads are a smaller slab).  0/1: one (default: measured faster than 6 or 8
                            passes even at 2^18 columns, where a tile's slab of B is 268 MB), k > 1: k */
  int32_t wide_tiles;    /* column tiles of B one work item covers: every pipeline stage then carries that many
                            B panels for ONE set of A images and a super-row has 512 / wide_tiles accumulator
                            columns.  0: the library's choice from n_hint (2 when n_hint spans >= 6 tile widths, or
                            >= 2 and the block-rows of a super-row rarely share a column block, else 1; measured,
                            DESIGN.md section 5), 1 / 2 / 4: forced.  Schedules with bounded accumulation chains
                            (tf32 by default) always run one tile per item. */
  int32_t n_hint;        /* expected number of B columns (0: unknown); only used to choose wide_tiles at create
                            time -- the one-shot calls pass their n */
} sparta_options;

/* Statistics of a handle (all counts refer to the handle's shard). */
typedef struct sparta_stats {
  int64_t rows;            /* C rows of the shard */
  int64_t cols;            /* A columns = B rows */
  int64_t block_rows;      /* block-rows in the shard */
  int64_t nz_blocks;       /* nonzero blocks */
  int64_t nztot;           /* sum h*w over nonzero blocks (VBR::nztot, vbr.cpp:232) */
  int64_t segments, super_rows, chunks, items;
  int64_t a_packed_bytes;  /* device bytes of the packed A images */
  int64_t b_bytes;         /* device bytes of the converted B operand */
  int64_t c_bytes;         /* device bytes of C */
  int32_t grid;            /* CTAs launched by sparta_run (2 per worker in pair mode) */
  int32_t smem_bytes;      /* dynamic shared memory per CTA */
  double  sched_imbalance; /* modelled max/mean CTA load */
  double  upload_ms;       /* host->device + packing time of the last create/set_B */
  int64_t kernel_launches; /* sm_100a SpMM launches issued through this handle */
  int32_t team;            /* workers walking one super-row side by side (column tiles per L2 pass) */
  int32_t cta_pair;        /* 1 when the handle runs CTA pairs */
  int32_t split_pieces;    /* (piece, column tile) items that add partial sums to C (split_k) */
  int32_t zero_tiles;      /* C tiles zeroed before each launch for those pieces */
  double  sched_max_cycles; /* modelled SM cycles of the worker that finishes last */
  int64_t gather_rows;     /* rows of the block-rows routed to the gather kernel (gather_max_height) */
  int64_t gather_nnz;      /* their nonzeros */
  int32_t wide_tiles;      /* column tiles of B per work item the handle was built for (1, 2 or 4) */
  int32_t reserved3;
} sparta_stats;

const char* sparta_last_error(void);
int sparta_abi_version(void);
/* number of visible CUDA devices with compute capability 10.x (0 on a CPU box) */
int sparta_device_count(void);

/* ---- handle API: device state persists across warm-up and repetitions ---- */

/* A in the reference's VBR layout (struct VBR, include/matrices.h:93-122):
 * row_part[block_rows+1], nzcount[block_rows], jab[sum nzcount] ascending per
 * block-row, mab = nonzero blocks back to back, each column-major with ld = h.
 * Host arrays are read during the call and never retained.
 * Replaces the upload half of cublas_fixed_blocks_multiply (cuda_utilities.cpp:91-124). */
int sparta_vbr_create(sparta_handle** out, int64_t rows, int64_t cols, int64_t block_rows,
                      int64_t block_col_size, const int64_t* row_part, const int64_t* nzcount,
                      const int64_t* jab, const float* mab, const sparta_options* opt);

/* A straight from the flat CSR (rowptr[rows+1], colind ascending per row, val or NULL for a pattern-only
 * matrix) and the row grouping (one group id per row, BlockingEngine::GetGrouping): replaces
 * VBR::fill_from_CSR_inplace (src/general/vbr.cpp:135-237) AND the upload half of the multiply routines.
 * block_col_size / row_block_size / force_fixed_size carry the meaning of the reference's call
 * (test/cuda/cuda_multiply.cpp:129, -b / -B / -F).  The host builds only the index arrays (identical to
 * the reference's row_part / nzcount / jab); the dense blocks are rebuilt on the device from the
 * nonzeros, so the fp32 mab -- 4.45 GB at BASELINE config #3 for 3.5 M nonzeros -- never exists on the
 * host or crosses PCIe.  C rows come back in blocked order like every VBR path; dims (may be NULL)
 * receives rows, cols, block_rows, block_cols, block_col_size, nztot of the VBR (rows / cols padded
 * when force_fixed_size).  The options' block_row_begin / block_row_end select block-rows. */
int sparta_vbr_create_from_csr(sparta_handle** out, int64_t rows, int64_t cols, const int64_t* rowptr,
                               const int64_t* colind, const float* val, const int64_t* grouping,
                               int64_t block_col_size, int64_t row_block_size, int32_t force_fixed_size,
                               const sparta_options* opt, int64_t* dims);

/* The INVERTED product C = B*A (-M 6 / -M 11): same VBR arrays, but B is n x rows and C is
 * n x cols, both column-major with ld >= n (cuda_utilities.cpp:556-559,587-591), i.e. the handle
 * computes C^T = A^T * B^T with the transposed blocks as the sparse operand: block-rows of the
 * operand are A's column blocks, the contraction runs over A's rows in blocked order, block
 * heights of any size become k extents.  set_B / get_C therefore take the ROW_MAJOR defaults
 * ([rows][n] and [cols][n] with ld >= n).  The options' block_row_begin / block_row_end select a
 * range of COLUMN BLOCKS (a slab of C's columns); the whole mab is uploaded for any range.
 * Intended arithmetic of cublas_blockmat_multiplyBA (cuda_utilities.cpp:640-690), whose own
 * indexing only works for constant heights and offsets B by block_col_size*ib (:645): see
 * DESIGN.md.  Replaces the upload half of that routine and of cutlas_blockmat_multiplyBA
 * (cutlass_bellpack_lib.cu:542-687). */
int sparta_vbr_create_BA(sparta_handle** out, int64_t rows, int64_t cols, int64_t block_rows,
                         int64_t block_col_size, const int64_t* row_part, const int64_t* nzcount,
                         const int64_t* jab, const float* mab, const sparta_options* opt);

/* A as the Blocked-ELL bundle produced by prepare_cusparse_BLOCKEDELLPACK
 * (cuda_utilities.cpp:1656-1710): ellColInd[ellColInd_rows*ellColInd_cols] with -1
 * padding, ellValues row-major rows x (ellColInd_cols*ell_blocksize).
 * Replaces the upload half of cusparse_gemm_custom_ellpack (:1497-1653) and
 * compute_cutlass_bellpack (cutlass_bellpack_lib.cu:61-242). */
int sparta_bellpack_create(sparta_handle** out, int64_t rows, int64_t cols,
                           int64_t ell_blocksize, int64_t ellColInd_rows,
                           int64_t ellColInd_cols, const int64_t* ellColInd,
                           const float* ellValues, const sparta_options* opt);

/* A as a flat CSR (rowptr[rows+1], colind ascending inside a row, val or NULL for a pattern-only
 * matrix whose entries are all 1) -- what prepare_cusparse_CSR (cuda_utilities.cpp:1433-1477)
 * flattens the reference's CSR struct into, with int64 indices.  B and C default to ROW_MAJOR like
 * the reference's cuSPARSE call (:1346-1355).  The options' block_row_begin / block_row_end select
 * a range of ROWS.  SPARTA_TF32 selects plain fp32 arithmetic here (no tensor cores on this
 * path): that mode is bit-identical to CSR::multiply (src/general/csr.cpp:49-65) for rows of up to 512
 * nonzeros; longer rows are summed as 8 slices whose partial sums are added in order (same products, a
 * different fp32 association: equal to the reference's sum within a few ulp, exactly for integer data).
 * Replaces the upload half of cusparse_gemm_custom (cuda_utilities.cpp:1251-1431, -M 2). */
int sparta_csr_create(sparta_handle** out, int64_t rows, int64_t cols, const int64_t* rowptr,
                      const int64_t* colind, const float* val, const sparta_options* opt);

/* Dense operand B (cols x n fp32).  ld: leading dimension in elements (>= cols for
 * COL_MAJOR, >= n for ROW_MAJOR).  on_device != 0: B is a device pointer on the
 * handle's device (e.g. the target of an NCCL broadcast). */
int sparta_set_B(sparta_handle* h, const float* B, int64_t ld, int64_t n, int on_device);

/* Initial C for handles created with accumulate = 1 (same layout/ld rules as sparta_get_C);
 * SPARTA_ERR_STATE on an accumulate = 0 handle, whose C is defined by the multiply alone. */
int sparta_set_C(sparta_handle* h, const float* C, int64_t ld, int on_device);

/* One multiply on the handle's stream.  *dt_ms (may be NULL) = CUDA-event time
 * around the compute kernel only, like the reference's dt (cuda_utilities.cpp:139,186-189). */
int sparta_run(sparta_handle* h, float* dt_ms);
/* Enqueue without timing or synchronisation (for external event timing / stream capture).
 * A handle whose plan has split pieces (sparta_stats.split_pieces > 0) passes a per-launch counter
 * target to the kernel, so a CAPTURED launch of it must not be replayed: create the handle with
 * split_k = 1 if the launch is to live in a CUDA graph. */
int sparta_run_async(sparta_handle* h);
int sparta_synchronize(sparta_handle* h);

/* Diagnostic: one multiply with the timeline of one worker (a CTA, or a CTA pair) recorded in SM
 * clock cycles.  records receives uint64[4][2][capacity][2] = zone x pair rank x index x {t0, t1}:
 * zone 0 producer per chunk {entered, TMA issued}; zone 1 MMA issuer per chunk, rank 0 {own stage
 * seen full, peer stage seen full}, rank 1 {peer stage seen full, MMAs issued}; zone 2 epilogue
 * per item {accumulator ready, drained}; zone 3 MMA issuer per item {waiting for the accumulator,
 * got it}.  Entries beyond what the worker executed stay zero. */
int sparta_run_traced(sparta_handle* h, int32_t worker, uint64_t* records, int64_t capacity);

/* Copy the result (rows x n fp32) out.  on_device != 0: C is a device pointer. */
int sparta_get_C(sparta_handle* h, float* C, int64_t ld, int on_device);

/* The result with its rows back in the ORIGINAL order: row r of the handle's C (blocked order,
 * vbr.cpp:355) is written to row row_map[r] of C, row_map = the reference's get_permutation
 * (src/general/utilities.cpp:8-20, sparta_host_permutation) restricted to the handle's rows.  C
 * has out_rows rows in the handle's C layout.  The reference leaves C in blocked order
 * (SURVEY.md 8a quirk 2); this is the un-permuting read-back of 8(f)-4.  on_device != 0: C is a
 * device buffer and only the mapped rows are written -- ranks of a multi-GPU run can scatter
 * their slabs straight into one full-size matrix.  Host C: unmapped rows come back as zeros. */
int sparta_get_C_permuted(sparta_handle* h, float* C, int64_t ld, const int64_t* row_map,
                          int64_t out_rows, int on_device);

/* Raw device views for zero-copy consumers (valid until the next set_B / destroy). */
void* sparta_C_device_ptr(sparta_handle* h);
int64_t sparta_C_device_ld(sparta_handle* h);
/* the CUDA stream (cudaStream_t) the handle launches on */
void* sparta_stream(sparta_handle* h);

int sparta_get_stats(sparta_handle* h, sparta_stats* out);
int sparta_destroy(sparta_handle* h);

/* ---- one-shot calls with the reference's data flow (host in, host out) ---- */

/* Replaces cublas_fixed_blocks_multiply (cuda_utilities.cpp:39-209, -M 4),
 * cublas_blockmat_batched (:723-887, -M 7) and the undefined
 * cublas_blockmat_multiplyAB (include/cuda_utilities.h:44): upload, multiply,
 * download.  B column-major ld = ldb, C column-major ld = ldc. */
int sparta_vbr_spmm(int64_t rows, int64_t cols, int64_t block_rows, int64_t block_col_size,
                    const int64_t* row_part, const int64_t* nzcount, const int64_t* jab,
                    const float* mab, const float* B, int64_t ldb, int64_t n, float* C,
                    int64_t ldc, int precision, float* dt_ms);

/* sparta_vbr_spmm on the first n_gpus sm_100 devices of the box (one process): A sharded by contiguous
 * block-row ranges balanced on modelled shard time, B uploaded ONCE and replicated by a single
 * ncclBroadcast over NVLink, no exchange during the multiply, C row-partitioned -- every device
 * copies its slab into the caller's C; gather_c != 0 all-gathers the slabs over NCCL first and
 * returns the whole C from device 0.  *dt_ms = the slowest device's kernel time, *bcast_ms (may be
 * NULL) = upload + broadcast of B.  NCCL is loaded at run time (libnccl.so.2); without it the call
 * fails with SPARTA_ERR_STATE.  The reference has no counterpart (SURVEY.md section 5, 8b "n_gpus,
 * gather_c"); integration/cuda_utilities_b200.cpp routes -M 4 / -M 7 here when SPARTA_GPUS > 1. */
int sparta_vbr_spmm_multi(int64_t rows, int64_t cols, int64_t block_rows, int64_t block_col_size,
                          const int64_t* row_part, const int64_t* nzcount, const int64_t* jab,
                          const float* mab, const float* B, int64_t ldb, int64_t n, float* C, int64_t ldc,
                          int precision, int32_t n_gpus, int32_t gather_c, float* dt_ms, float* bcast_ms);

/* The same one-shot flow from the CSR and the grouping (sparta_vbr_create_from_csr): what
 * fill_from_CSR_inplace + cublas_fixed_blocks_multiply do together in cuda_multiply.cpp:129-137.
 * B column-major ld = ldb >= cols (padded cols when force_fixed_size), C column-major ld = ldc >= rows. */
int sparta_csr_vbr_spmm(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind, const float* val,
                        const int64_t* grouping, int64_t block_col_size, int64_t row_block_size,
                        int32_t force_fixed_size, const float* B, int64_t ldb, int64_t n, float* C, int64_t ldc,
                        int precision, float* dt_ms);

/* Replaces cublas_blockmat_multiplyBA (cuda_utilities.cpp:553-721, -M 6) and
 * cutlas_blockmat_multiplyBA (cutlass_bellpack_lib.cu:542, -M 11): C (n x cols) = B (n x rows) * A,
 * B and C column-major with ldb, ldc >= n. */
int sparta_vbr_spmm_BA(int64_t rows, int64_t cols, int64_t block_rows, int64_t block_col_size,
                       const int64_t* row_part, const int64_t* nzcount, const int64_t* jab,
                       const float* mab, const float* B, int64_t ldb, int64_t n, float* C,
                       int64_t ldc, int precision, float* dt_ms);

/* Replaces cusparse_gemm_custom_ellpack (-M 3) / compute_cutlass_bellpack (-M 8).
 * B and C row-major (cuda_utilities.cpp:1581-1591). */
int sparta_bellpack_spmm(int64_t rows, int64_t cols, int64_t ell_blocksize,
                         int64_t ellColInd_rows, int64_t ellColInd_cols,
                         const int64_t* ellColInd, const float* ellValues, const float* B,
                         int64_t ldb, int64_t n, float* C, int64_t ldc, int precision,
                         float* dt_ms);

/* Replaces cusparse_blockmat_multiplyAB (cuda_utilities.cpp:1479-1493, -M 2): flat CSR in,
 * B and C row-major. */
int sparta_csr_spmm(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                    const float* val, const float* B, int64_t ldb, int64_t n, float* C, int64_t ldc,
                    int precision, float* dt_ms);

/* Device buffers are taken from the device's stream-ordered memory pool and stay cached there
 * after sparta_destroy / a one-shot call, so that repeated calls do not pay cudaMalloc / cudaFree
 * (the reference re-allocates on every call, cuda_utilities.cpp:95-98,204-206).  This returns the
 * cached memory of every device the library has used to the driver. */
int sparta_release_workspace(void);

/* ---- host-side helpers (no GPU needed) ---- */

/* Contiguous block-row ranges balanced on nonzero-block area; cuts[parts+1]. */
int sparta_partition_block_rows(int64_t block_rows, const int64_t* row_part,
                                const int64_t* nzcount, int32_t parts, int64_t* cuts);

/* The same kind of partition balanced on the MODELLED kernel time of every shard (the tile
 * scheduler's cost model: bytes staged per chunk, drain per work item) rather than on area: sparse
 * block-rows cost more per FLOP than dense ones.  n = columns of B, opt = the options the handles
 * will be created with (NULL for defaults). */
int sparta_partition_block_rows_modelled(int64_t rows, int64_t cols, int64_t block_rows,
                                         int64_t block_col_size, const int64_t* row_part,
                                         const int64_t* nzcount, const int64_t* jab, int64_t n,
                                         const sparta_options* opt, int32_t parts, int64_t* cuts);

/* The model's time (SM cycles) of every shard of a GIVEN partition cuts[parts+1]: what
 * sparta_partition_block_rows_measured needs next to the measured times. */
int sparta_partition_model_times(int64_t rows, int64_t cols, int64_t block_rows, int64_t block_col_size,
                                 const int64_t* row_part, const int64_t* nzcount, const int64_t* jab, int64_t n,
                                 const sparta_options* opt, int32_t parts, const int64_t* cuts, double* cycles);

/* The modelled partition corrected by MEASURED times: time_scale[block_rows] holds, for every
 * block-row, measured / modelled kernel time of the shard it belonged to in an earlier partition
 * (ranks time a few launches, all-gather the times and re-cut). */
int sparta_partition_block_rows_measured(int64_t rows, int64_t cols, int64_t block_rows,
                                         int64_t block_col_size, const int64_t* row_part,
                                         const int64_t* nzcount, const int64_t* jab, int64_t n,
                                         const sparta_options* opt, int32_t parts,
                                         const double* time_scale, int64_t* cuts);

/* ---- host-side format builders (no GPU needed; bit-exact with the reference) ---- */

/* Row clustering: BlockingEngine::GetGrouping (src/general/blocking.cpp:633-676) on a flat CSR
 * with strictly ascending columns per row.  Parameters carry the reference CLI meaning
 * (include/input.h:15-42): algo = -a (0 iterative, 2 fixed, 3 clocked, 4 queue, 5 max-size),
 * tau = -t, block_col_size = -b, row_block_size = -B, sim_measure = -m (0 Hamming, 1 Jaccard,
 * 2/3 their probe variants), use_pattern = -p, use_groups = -g, force_fixed_size = -F.
 * grouping[rows] receives one group id per row, identical to the reference's.  stats (may be
 * NULL) receives the counters save_blocking_data prints (src/general/utilities.cpp:175-233).
 * flags bit 0: evaluate Jaccard through the column-list model instead of the block bitmap. */
typedef struct sparta_blocking_stats {
  int64_t comparison_counter, merge_counter;
  float average_merge_tau, average_row_distance;
  double seconds;
} sparta_blocking_stats;
int sparta_host_blocking(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                         int32_t algo, float tau, int64_t block_col_size, int64_t row_block_size,
                         int32_t sim_measure, int32_t use_pattern, int32_t use_groups,
                         int32_t force_fixed_size, int32_t flags, int64_t* grouping,
                         sparta_blocking_stats* stats);

/* Grouping files.  The reference persists a grouping as `<outfile>.g`, one group id per line
 * (test/general/Matrix_Blocking.cpp:24-32, src/general/utilities.cpp:240-243) and reloads such a file
 * in test/general/Matrix_Analysis.cpp:10-32.  sparta_grouping_save / _load read and write exactly that
 * format, plus a sidecar `<path>.key` (hex key, row count, flags as text): _load with key != 0 fails
 * unless the sidecar holds the same key.  sparta_blocking_key hashes the CSR pattern and the blocking
 * flags (0 on invalid input).  sparta_host_blocking_cached looks `<cache_dir>/grouping_<key>.g` up,
 * runs sparta_host_blocking and stores the result on a miss; *hit (may be NULL) tells which happened.
 * Blocking a 2^18-row R-MAT with -a 4 takes 17 minutes of CPU; the file is 1.6 MB. */
int sparta_grouping_save(const char* path, int64_t rows, const int64_t* grouping, uint64_t key);
int sparta_grouping_load(const char* path, int64_t rows, int64_t* grouping, uint64_t key);
uint64_t sparta_blocking_key(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                             int32_t algo, float tau, int64_t block_col_size, int64_t row_block_size,
                             int32_t sim_measure, int32_t use_pattern, int32_t use_groups,
                             int32_t force_fixed_size);
int sparta_host_blocking_cached(const char* cache_dir, int64_t rows, int64_t cols, const int64_t* rowptr,
                                const int64_t* colind, int32_t algo, float tau, int64_t block_col_size,
                                int64_t row_block_size, int32_t sim_measure, int32_t use_pattern,
                                int32_t use_groups, int32_t force_fixed_size, int32_t flags,
                                int64_t* grouping, sparta_blocking_stats* stats, int32_t* hit);

/* The reference's -r row reorderings applied right after reading (include/matrices.h:65-82,
 * src/general/csr.cpp:123-166): order[i] = the old index of new row i.  mode -1: ascending degree
 * (CSR::reorder_by_degree), 2: CSR::scramble = std::random_shuffle driven by std::rand seeded from -s
 * (include/input.h:111-114; seed 0 leaves the generator alone) -- the same glibc generator stepped the way
 * libstdc++'s random_shuffle steps it, hence the same permutation.  mode 1 (descending degree) is refused:
 * the reference hands std::sort a >= comparator (csr.cpp:130-133), which is undefined behaviour. */
int sparta_host_row_order(int64_t rows, const int64_t* rowptr, int32_t mode, uint32_t seed, int64_t* order);

/* get_permutation / get_partition (src/general/utilities.cpp:8-43).  perm[n]; part needs
 * n+1 slots, *part_len receives block_rows+1. */
int sparta_host_permutation(int64_t n, const int64_t* grouping, int64_t* perm);
int sparta_host_partition(int64_t n, const int64_t* grouping, int64_t* part, int64_t* part_len);

/* VBR::fill_from_CSR_inplace (src/general/vbr.cpp:135-237) in linear time on a flat CSR
 * (rowptr[rows+1], colind, val; val may be NULL when pattern_only).  The result is owned by
 * the returned object; sparta_host_vbr_get exposes its arrays (valid until _free).
 * dims[6] = rows, cols, block_rows, block_cols, block_col_size, nztot. */
typedef struct sparta_host_vbr sparta_host_vbr;
int sparta_host_vbr_fill(sparta_host_vbr** out, int64_t rows, int64_t cols, const int64_t* rowptr,
                         const int64_t* colind, const float* val, int32_t pattern_only,
                         const int64_t* grouping, int64_t block_col_size, int64_t row_block_size,
                         int32_t force_fixed_size, int32_t threads);
int sparta_host_vbr_get(sparta_host_vbr* v, int64_t* dims, const int64_t** row_part,
                        const int64_t** nzcount, const int64_t** jab, const float** mab);
int sparta_host_vbr_free(sparta_host_vbr* v);

/* prepare_cusparse_BLOCKEDELLPACK (src/cuda/cuda_utilities.cpp:1656-1710).
 * dims[3] = ell_blocksize, ellColInd_rows, ellColInd_cols. */
typedef struct sparta_host_bell sparta_host_bell;
int sparta_host_bellpack_from_vbr(sparta_host_bell** out, int64_t rows, int64_t cols,
                                  int64_t block_col_size, const int64_t* nzcount,
                                  const int64_t* jab, const float* mab, int32_t threads);
int sparta_host_bellpack_get(sparta_host_bell* b, int64_t* dims, const int64_t** ellColInd,
                             const float** ellValues);
int sparta_host_bellpack_free(sparta_host_bell* b);

/* Host-only access to the schedule itself so the CPU test-suite can interpret it
 * (tests/sched_interp.py) without a GPU.  No arithmetic on matrix values happens
 * in the library on this path.  `which`: 0 segments, 1 super-rows, 2 chunks,
 * 3 items, 4 cta_ptr, 5 cta_items, 6 pack jobs, 7 run tables, 8 zero jobs (record layouts:
 * csrc/sched_types.h).
 * *data points into the plan and stays valid until sparta_plan_destroy. */
typedef struct sparta_plan sparta_plan;
int sparta_vbr_plan_create(sparta_plan** out, int64_t rows, int64_t cols, int64_t block_rows,
                           int64_t block_col_size, const int64_t* row_part,
                           const int64_t* nzcount, const int64_t* jab, int64_t n,
                           const sparta_options* opt);
/* the schedule of the inverted product (sparta_vbr_create_BA) */
int sparta_vbr_plan_create_BA(sparta_plan** out, int64_t rows, int64_t cols, int64_t block_rows,
                              int64_t block_col_size, const int64_t* row_part,
                              const int64_t* nzcount, const int64_t* jab, int64_t n,
                              const sparta_options* opt);
int sparta_plan_array(sparta_plan* plan, int32_t which, const void** data, int64_t* count,
                      int32_t* record_bytes);
int sparta_plan_stats(sparta_plan* plan, sparta_stats* out);
int sparta_plan_destroy(sparta_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* SPARTA_B200_H */
