// Drop-in replacement for the reference's GPU multiplication layer.
//
// Compile THIS file in place of src/cuda/cuda_utilities.cpp and
// src/cuda/cutlass_bellpack_lib.cu inside the SPARTA tree (it includes the reference's own
// headers, nothing from the reference is copied here) and link libsparta_b200.so instead of
// -lcublas -lcusparse and CUTLASS.  test/cuda/cuda_multiply.cpp then builds UNCHANGED and its
// -M switch (include/definitions.h:19) runs on the sm_100a kernel family:
//
//   -M 2  cusparse_spmm          cusparse_blockmat_multiplyAB      -> sparta_csr_spmm      (fp32)
//   -M 3  cusparse_bellpack      bellpack_blockmat_multiplyAB      -> sparta_bellpack_spmm (fp16)
//   -M 4  cublas_vbr             cublas_fixed_blocks_multiply      -> sparta_vbr_spmm      (fp16)
//   -M 6  cublas_vbr_inverted    cublas_blockmat_multiplyBA        -> sparta_vbr_spmm_BA   (fp16)
//   -M 7  cublas_vbr_batched     cublas_blockmat_batched           -> sparta_vbr_spmm      (tf32)
//   -M 8  cutlass_bellpack       bellpack_cutlass_multiplyAB       -> sparta_bellpack_spmm (fp16)
//   -M 10 cutlas_vbr             cutlas_fixed_blocks_multiply      -> sparta_vbr_spmm      (fp16)
//   -M 11 cutlas_vbr_inverted    cutlas_blockmat_multiplyBA        -> sparta_vbr_spmm_BA   (fp16)
//   (undefined in the reference) cublas_blockmat_multiplyAB        -> sparta_vbr_spmm, true variable heights
// and test/cuda/TEST_cuda.cpp (prepare_cusparse_CSR, cusparse_gemm_custom, the pico_print_SpMMM printers)
// builds unchanged too.  SPARTA_B200_CSV=<file> adds the result columns the reference's CSV lacks.
//
// Operand precision follows what each reference routine asks of its library (CUDA_R_16F for the
// cuBLAS/cuSPARSE/CUTLASS paths, cuda_utilities.cpp:29-31; fp32 SGEMM for the batched path,
// :859) and can be overridden with SPARTA_PRECISION=bf16|fp16|tf32.  Semantics kept: host
// pointers in and out, alpha = beta = 1 on a caller-zeroed C, dt = CUDA-event milliseconds around
// the compute only, any failure prints and exits like checkCudaErrors (helper_cuda.h:714-727).
//
// Not provided (outside the hot path): the batched inverted product (-M 12) and the dense GEMMs
// (-M 1, 9); they print a message and exit.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cuda_utilities.h"
#include "cutlass_bellpack_lib.h"

#include "sparta_b200.h"

namespace {

[[noreturn]] void die(const char* what) {
  fprintf(stderr, "libsparta_b200: %s: %s\n", what, sparta_last_error());
  exit(EXIT_FAILURE);
}

[[noreturn]] void not_provided(const char* fn, const char* flag) {
  fprintf(stderr, "%s (%s) is outside the sm_100a hot path and is not provided by this build\n", fn, flag);
  exit(EXIT_FAILURE);
}

int precision_or(int fallback) {
  const char* p = getenv("SPARTA_PRECISION");
  if (!p) return fallback;
  if (!strcmp(p, "bf16")) return SPARTA_BF16;
  if (!strcmp(p, "fp16")) return SPARTA_FP16;
  if (!strcmp(p, "tf32")) return SPARTA_TF32;
  fprintf(stderr, "SPARTA_PRECISION must be bf16, fp16 or tf32\n");
  exit(EXIT_FAILURE);
}

// ---- result columns the reference's CSV does not have --------------------------------
// save_blocking_data (src/general/utilities.cpp:175-233) writes one row per run with the average
// kernel time only, and the CLI must stay untouched.  With SPARTA_B200_CSV=<file> every multiply
// routine of this shim appends one row to that side file: routine, shape, nonzero blocks, nztot,
// B columns, precision, GPUs, dt, effective TFLOP/s on nonzero-block FLOPs (2*nztot*n/dt, the
// reference's own quantity times two, src/scripts/multiplication_barplots.py:515), algorithmic bytes,
// GB/s, and the fractions of the tensor and HBM peaks (SPARTA_PEAK_TFLOPS / SPARTA_PEAK_GBS, defaults:
// the B200 figures this repo measured, 1615.4 bf16 TFLOP/s -- half for tf32 -- and 6547.5 GB/s).
void metrics_row(const char* routine, long rows, long cols, long nz_blocks, double nztot, int n, int precision,
                 float dt_ms, double algorithmic_bytes) {
  const char* path = getenv("SPARTA_B200_CSV");
  if (!path || !*path) return;
  FILE* probe = fopen(path, "r");
  const bool fresh = probe == nullptr;
  if (probe) fclose(probe);
  FILE* f = fopen(path, "a");
  if (!f) return;
  if (fresh)
    fprintf(f, "routine,rows,cols,nz_blocks,nztot,b_cols,precision,gpus,dt_ms,effective_tflops,algorithmic_bytes,"
               "algorithmic_gbs,pct_tensor_peak,pct_hbm_peak\n");
  const char* pt = getenv("SPARTA_PEAK_TFLOPS");
  const char* pg = getenv("SPARTA_PEAK_GBS");
  double peak_t = pt ? atof(pt) : 1615.4;
  if (!pt && precision == SPARTA_TF32) peak_t *= 0.5;
  const double peak_g = pg ? atof(pg) : 6547.5;
  const double sec = dt_ms * 1e-3;
  const double tflops = sec > 0 ? 2.0 * nztot * n / sec / 1e12 : 0.0;
  const double gbs = sec > 0 ? algorithmic_bytes / sec / 1e9 : 0.0;
  static const char* names[3] = {"bf16", "fp16", "tf32"};
  fprintf(f, "%s,%ld,%ld,%ld,%.0f,%d,%s,1,%.6f,%.3f,%.0f,%.1f,%.2f,%.2f\n", routine, rows, cols, nz_blocks, nztot, n,
          names[precision], static_cast<double>(dt_ms), tflops, algorithmic_bytes, gbs, 100.0 * tflops / peak_t,
          100.0 * gbs / peak_g);
  fclose(f);
}

// A once + B once + C once, in the bytes the kernel moves (2 per operand element, 4 for tf32 and C)
double vbr_bytes(const VBR& A, double nztot, int n, int precision) {
  const double es = precision == SPARTA_TF32 ? 4 : 2;
  return nztot * es + static_cast<double>(A.cols) * n * es + static_cast<double>(A.rows) * n * 4;
}

double vbr_nztot(const VBR& A, long* blocks) {
  double t = 0;
  long nb = 0;
  for (intT ib = 0; ib < A.block_rows; ++ib) {
    t += static_cast<double>(A.nzcount[ib]) * (A.row_part[ib + 1] - A.row_part[ib]) * A.block_col_size;
    nb += A.nzcount[ib];
  }
  *blocks = nb;
  return t;
}

// B column-major (ld = A.cols), C column-major (ld = A.rows), like VBR::multiply (vbr.cpp:331,355)
void vbr_multiply(const VBR& A, DataT* B, int B_cols, DataT_C* C, float& dt, int precision) {
  static_assert(sizeof(intT) == sizeof(int64_t) && sizeof(DataT) == sizeof(float) && sizeof(DataT_C) == sizeof(float),
                "the C ABI takes the reference's default types (definitions.h:4-6)");
  // SPARTA_GPUS=k (k > 1): the same call on k GPUs of the box -- A sharded by block-rows, B replicated
  // by one NCCL broadcast, C row-partitioned (SPARTA_GATHER_C=1: all-gathered over NCCL first)
  const char* g = getenv("SPARTA_GPUS");
  const int gpus = g ? atoi(g) : 1;
  if (gpus > 1) {
    const char* gc = getenv("SPARTA_GATHER_C");
    float bcast_ms = 0;
    if (sparta_vbr_spmm_multi(A.rows, A.cols, A.block_rows, A.block_col_size,
                              reinterpret_cast<const int64_t*>(A.row_part), reinterpret_cast<const int64_t*>(A.nzcount),
                              reinterpret_cast<const int64_t*>(A.jab), A.mab, B, A.cols, B_cols, C, A.rows, precision,
                              gpus, gc && atoi(gc), &dt, &bcast_ms))
      die("sparta_vbr_spmm_multi");
  } else if (sparta_vbr_spmm(A.rows, A.cols, A.block_rows, A.block_col_size,
                             reinterpret_cast<const int64_t*>(A.row_part), reinterpret_cast<const int64_t*>(A.nzcount),
                             reinterpret_cast<const int64_t*>(A.jab), A.mab, B, A.cols, B_cols, C, A.rows, precision, &dt))
    die("sparta_vbr_spmm");
  long nb = 0;
  const double nztot = vbr_nztot(A, &nb);
  metrics_row("vbr_multiply", A.rows, A.cols, nb, nztot, B_cols, precision, dt, vbr_bytes(A, nztot, B_cols, precision));
}

// C = B*A: B is B_rows x A.rows, C is B_rows x A.cols, both column-major with ld = B_rows
// (cuda_utilities.cpp:556-559,587-591)
void vbr_multiply_BA(const VBR& A, DataT* B, int B_rows, DataT_C* C, float& dt, int precision) {
  if (sparta_vbr_spmm_BA(A.rows, A.cols, A.block_rows, A.block_col_size,
                         reinterpret_cast<const int64_t*>(A.row_part), reinterpret_cast<const int64_t*>(A.nzcount),
                         reinterpret_cast<const int64_t*>(A.jab), A.mab, B, B_rows, B_rows, C, B_rows, precision, &dt))
    die("sparta_vbr_spmm_BA");
  long nb = 0;
  const double nztot = vbr_nztot(A, &nb);
  const double es = precision == SPARTA_TF32 ? 4 : 2;
  metrics_row("vbr_multiply_BA", A.rows, A.cols, nb, nztot, B_rows, precision, dt,
              nztot * es + static_cast<double>(A.rows) * B_rows * es + static_cast<double>(A.cols) * B_rows * 4);
}

}  // namespace

// ---- VBR x dense ------------------------------------------------------------------------------

void cublas_fixed_blocks_multiply(const VBR& vbmatA, DataT* B, int B_cols, DataT_C* C, float& dt, int /*n_streams*/) {
  vbr_multiply(vbmatA, B, B_cols, C, dt, precision_or(SPARTA_FP16));
}

void cublas_blockmat_multiplyAB(const VBR& vbmatA, DataT* B, int B_cols, DataT_C* C, float& dt, int /*n_streams*/) {
  vbr_multiply(vbmatA, B, B_cols, C, dt, precision_or(SPARTA_FP16));
}

void cublas_blockmat_batched(const VBR& vbmatA, DataT* B, int B_cols, DataT_C* C, float& dt) {
  vbr_multiply(vbmatA, B, B_cols, C, dt, precision_or(SPARTA_TF32));
}

void cutlas_fixed_blocks_multiply(const VBR& vbmatA, DataT* B, int B_cols, DataT_C* C, float& dt) {
  vbr_multiply(vbmatA, B, B_cols, C, dt, precision_or(SPARTA_FP16));
}

// ---- dense x VBR (the inverted product, -M 6 / -M 11) ------------------------------------------

void cublas_blockmat_multiplyBA(const VBR& vbmatA, DataT* B, int B_rows, DataT_C* C, float& dt, int /*n_streams*/) {
  vbr_multiply_BA(vbmatA, B, B_rows, C, dt, precision_or(SPARTA_FP16));
}

void cutlas_blockmat_multiplyBA(const VBR& vbmatA, DataT* B, int B_rows, DataT_C* C, float& dt) {
  vbr_multiply_BA(vbmatA, B, B_rows, C, dt, precision_or(SPARTA_FP16));
}

void cutlas_blockmat_multiplyBA_streams(const VBR& vbmatA, DataT* B, int B_rows, DataT_C* C, float& dt, int /*n_streams*/) {
  vbr_multiply_BA(vbmatA, B, B_rows, C, dt, precision_or(SPARTA_FP16));
}

// ---- Blocked-ELL x dense (B and C row-major, cuda_utilities.cpp:1581-1591) ---------------------

int prepare_cusparse_BLOCKEDELLPACK(VBR* A, int* ell_blocksize, int* ellValue_cols, int* ellColInd_rows,
                                    int* ellColInd_cols, int* num_blocks, intT** ellColInd, DataT_C** ellValues) {
  sparta_host_bell* bell = nullptr;
  if (sparta_host_bellpack_from_vbr(&bell, A->rows, A->cols, A->block_col_size,
                                    reinterpret_cast<const int64_t*>(A->nzcount),
                                    reinterpret_cast<const int64_t*>(A->jab), A->mab, 0)) {
    fprintf(stderr, "prepare_cusparse_BLOCKEDELLPACK: %s\n", sparta_last_error());
    exit(__LINE__);   // the reference exits with a line number here (cuda_utilities.cpp:1664-1670)
  }
  int64_t dims[3];
  const int64_t* ind = nullptr;
  const float* vals = nullptr;
  sparta_host_bellpack_get(bell, dims, &ind, &vals);
  *ell_blocksize = static_cast<int>(dims[0]);
  *ellColInd_rows = static_cast<int>(dims[1]);
  *ellColInd_cols = static_cast<int>(dims[2]);
  *ellValue_cols = static_cast<int>(dims[2] * dims[0]);
  *num_blocks = static_cast<int>(dims[1] * dims[2]);
  const size_t n_ind = static_cast<size_t>(dims[1]) * dims[2];
  const size_t n_val = static_cast<size_t>(A->rows) * (*ellValue_cols);
  *ellColInd = new intT[n_ind ? n_ind : 1];     // the caller owns both arrays, as in the reference
  *ellValues = new DataT_C[n_val ? n_val : 1];
  for (size_t i = 0; i < n_ind; ++i) (*ellColInd)[i] = static_cast<intT>(ind[i]);
  if (n_val) memcpy(*ellValues, vals, n_val * sizeof(DataT_C));
  sparta_host_bellpack_free(bell);
  return 0;
}

int cusparse_gemm_custom_ellpack(int rows, int cols, int A_ell_blocksize, int /*A_ellValues_cols*/, int A_ellColInd_cols,
                                 int A_ellColInd_rows, int /*A_num_blocks*/, intT* A_ellColInd, DataT_C* A_ellValues,
                                 DataT* B, int B_cols, int B_lead_dim, DataT_C* C, int C_lead_dim,
                                 const DataT_C /*alpha*/, const DataT_C /*beta*/, float& dt) {
  if (sparta_bellpack_spmm(rows, cols, A_ell_blocksize, A_ellColInd_rows, A_ellColInd_cols,
                           reinterpret_cast<const int64_t*>(A_ellColInd), A_ellValues, B, B_lead_dim, B_cols, C,
                           C_lead_dim, precision_or(SPARTA_FP16), &dt))
    die("sparta_bellpack_spmm");
  long real_blocks = 0;
  for (long i = 0; i < static_cast<long>(A_ellColInd_rows) * A_ellColInd_cols; ++i) real_blocks += A_ellColInd[i] >= 0;
  const double nztot = static_cast<double>(real_blocks) * A_ell_blocksize * A_ell_blocksize;
  const int prec = precision_or(SPARTA_FP16);
  const double es = prec == SPARTA_TF32 ? 4 : 2;
  metrics_row("cusparse_gemm_custom_ellpack", rows, cols, real_blocks, nztot, B_cols, prec, dt,
              nztot * es + static_cast<double>(cols) * B_cols * es + static_cast<double>(rows) * B_cols * 4);
  return 0;
}

void bellpack_blockmat_multiplyAB(VBR* A, DataT* B, int B_cols, DataT_C* C, int C_cols, float& dt, int /*verbose*/) {
  int bs, val_cols, ind_rows, ind_cols, num_blocks;
  intT* ind = nullptr;
  DataT_C* vals = nullptr;
  prepare_cusparse_BLOCKEDELLPACK(A, &bs, &val_cols, &ind_rows, &ind_cols, &num_blocks, &ind, &vals);
  cusparse_gemm_custom_ellpack(A->rows, A->cols, bs, val_cols, ind_cols, ind_rows, num_blocks, ind, vals, B, B_cols,
                               B_cols, C, C_cols, 1, 1, dt);
  delete[] ind;
  delete[] vals;
}

void bellpack_cutlass_multiplyAB(VBR* A, DataT* B, int B_cols, DataT_C* C, int C_cols, float& dt, int verbose) {
  bellpack_blockmat_multiplyAB(A, B, B_cols, C, C_cols, dt, verbose);
}

// ---- CSR x dense (B and C row-major, cuda_utilities.cpp:1346-1355) ------------------------------

// prepare_cusparse_CSR (cuda_utilities.cpp:1433-1477): the reference's array-of-rows CSR flattened
// into malloc'ed int32 rowptr / colind and a value array (ones for a pattern-only matrix); the caller
// frees the three arrays with free(), as cusparse_blockmat_multiplyAB does (:1487-1489).
int prepare_cusparse_CSR(CSR& cmat, int** csrRowPtr, int** csrColInd, DataT** csrVal) {
  intT nnz = 0;
  for (intT i = 0; i < cmat.rows; ++i) nnz += cmat.nzcount[i];
  *csrRowPtr = static_cast<int*>(malloc((static_cast<size_t>(cmat.rows) + 1) * sizeof(int)));
  *csrColInd = static_cast<int*>(malloc((nnz ? nnz : 1) * sizeof(int)));
  *csrVal = static_cast<DataT*>(malloc((nnz ? nnz : 1) * sizeof(DataT)));
  if (!*csrRowPtr || !*csrColInd || !*csrVal) { fprintf(stderr, "prepare_cusparse_CSR: out of memory\n"); exit(EXIT_FAILURE); }
  intT at = 0;
  for (intT i = 0; i < cmat.rows; ++i) {
    (*csrRowPtr)[i] = static_cast<int>(at);
    for (intT q = 0; q < cmat.nzcount[i]; ++q, ++at) {
      (*csrColInd)[at] = static_cast<int>(cmat.ja[i][q]);
      (*csrVal)[at] = cmat.pattern_only ? DataT(1) : cmat.ma[i][q];
    }
  }
  (*csrRowPtr)[cmat.rows] = static_cast<int>(at);
  return 0;
}

// cusparse_gemm_custom (cuda_utilities.cpp:1251-1431): int32 CSR x dense, B and C row-major with
// the given leading dimensions.  The reference asks cuSPARSE for CUDA_R_32F compute (:1267), hence
// the fp32 default (bit-identical to CSR::multiply); alpha = beta = 1 on a caller-zeroed C like every
// call site (cuda_multiply.cpp:276, TEST_cuda.cpp:173).
int cusparse_gemm_custom(int rows, int cols, int nnz, int* csrRowPtr, int* csrColInd, DataT* csrVal, DataT* B,
                         int B_cols, int B_lead_dim, DataT_C* C, int C_lead_dim, const DataT_C alpha,
                         const DataT_C beta, float& dt) {
  if (alpha != DataT_C(1) || beta != DataT_C(1)) {
    fprintf(stderr, "cusparse_gemm_custom: only alpha = beta = 1 (every call site of the reference) is provided\n");
    exit(EXIT_FAILURE);
  }
  std::vector<int64_t> rowptr(csrRowPtr, csrRowPtr + rows + 1), colind(csrColInd, csrColInd + nnz);
  const int prec = precision_or(SPARTA_TF32);
  if (sparta_csr_spmm(rows, cols, rowptr.data(), colind.data(), csrVal, B, B_lead_dim, B_cols, C, C_lead_dim, prec, &dt))
    die("sparta_csr_spmm");
  metrics_row("cusparse_gemm_custom", rows, cols, nnz, nnz, B_cols, prec, dt,
              static_cast<double>(nnz) * 8 + static_cast<double>(cols) * B_cols * 4 + static_cast<double>(rows) * B_cols * 4);
  return 0;
}

void cusparse_blockmat_multiplyAB(CSR& A, DataT* B, int B_cols, DataT_C* C, int C_cols, float& dt) {
  DataT* csrVal;
  int *csrRowPtr, *csrColInd;
  prepare_cusparse_CSR(A, &csrRowPtr, &csrColInd, &csrVal);
  cusparse_gemm_custom(A.rows, A.cols, static_cast<int>(A.nztot()), csrRowPtr, csrColInd, csrVal, B, B_cols, B_cols, C,
                       C_cols, 1, 1, dt);
  free(csrVal);
  free(csrColInd);
  free(csrRowPtr);
}

// ---- debug printers the CLI calls at -v 3 ------------------------------------------------------

void pico_print_DnM(const char* Cname, int Cn, int Cm, DataT_C* C) {
  printf("Dense matrix %s (%d x %d):\n", Cname, Cn, Cm);
  for (int i = 0; i < Cn; ++i) {
    for (int j = 0; j < Cm; ++j) printf("%g ", static_cast<double>(C[i * Cm + j]));
    printf("\n");
  }
}

static void print_dense(const char* name, long n, long m, const float* M) {
  if (!M || !strcmp(name, "NULL")) return;
  printf("%s (%ld x %ld, row-major):\n", name, n, m);
  for (long i = 0; i < n; ++i) {
    for (long j = 0; j < m; ++j) printf("%6.2f ", static_cast<double>(M[i * m + j]));
    printf("\n");
  }
}

// The three PICO_DEBUG printers of include/cuda_utilities.h:48-52 (test/cuda/TEST_cuda.cpp calls them
// when built with -DPICO_DEBUG): operand A in the given format, then B and C when they are named.
void pico_print_SpMMM(const char* Aname, int An, int Am, int Az, int* Arows, int* Acols, DataT* Avals,
                      const char* Bname, int Bn, int Bm, DataT* B, const char* Cname, long int Cn, long int Cm, DataT_C* C) {
  if (Arows && strcmp(Aname, "NULL")) {
    printf("%s: CSR %d x %d, %d nonzeros\n", Aname, An, Am, Az);
    for (int i = 0; i < An; ++i)
      for (int q = Arows[i]; q < Arows[i + 1]; ++q) printf("  (%d, %d) = %g\n", i, Acols[q], static_cast<double>(Avals[q]));
  }
  print_dense(Bname, Bn, Bm, B);
  print_dense(Cname, Cn, Cm, C);
}

void pico_print_SpMMM(const char* Aname, VBR* A, const char* Bname, int Bn, int Bm, DataT* B, const char* Cname,
                      long int Cn, long int Cm, DataT_C* C) {
  if (A && strcmp(Aname, "NULL")) {
    printf("%s: VBR %ld x %ld, %ld block-rows, column blocks of %ld, nztot %ld\n", Aname, static_cast<long>(A->rows),
           static_cast<long>(A->cols), static_cast<long>(A->block_rows), static_cast<long>(A->block_col_size),
           static_cast<long>(A->nztot));
    const intT* jab = A->jab;
    for (intT ib = 0; ib < A->block_rows; ++ib) {
      printf("  block-row %ld rows [%ld, %ld):", static_cast<long>(ib), static_cast<long>(A->row_part[ib]),
             static_cast<long>(A->row_part[ib + 1]));
      for (intT q = 0; q < A->nzcount[ib]; ++q) printf(" %ld", static_cast<long>(*jab++));
      printf("\n");
    }
  }
  print_dense(Bname, Bn, Bm, B);
  print_dense(Cname, Cn, Cm, C);
}

void pico_print_SpMMM(const char* Aname, int rows, int cols, int ell_blocksize, int ellValue_cols, int ellColumnsInd_rows,
                      int ellColumnsInd_cols, int num_blocks, intT* ellColumnsInd, DataT_C* ellValues, const char* Bname,
                      int Bn, int Bm, DataT* B, const char* Cname, long int Cn, long int Cm, DataT_C* C) {
  if (ellColumnsInd && strcmp(Aname, "NULL")) {
    printf("%s: Blocked-ELL %d x %d, block %d, %d x %d index grid (%d slots), %d value columns\n", Aname, rows, cols,
           ell_blocksize, ellColumnsInd_rows, ellColumnsInd_cols, num_blocks, ellValue_cols);
    for (int i = 0; i < ellColumnsInd_rows; ++i) {
      printf("  ");
      for (int j = 0; j < ellColumnsInd_cols; ++j) printf("%ld ", static_cast<long>(ellColumnsInd[i * ellColumnsInd_cols + j]));
      printf("\n");
    }
    print_dense("ellValues", rows, ellValue_cols, ellValues);
  }
  print_dense(Bname, Bn, Bm, B);
  print_dense(Cname, Cn, Cm, C);
}

// ---- outside the hot path ----------------------------------------------------------------------

void cutlas_blockmat_batched(const VBR&, DataT*, int, DataT_C*, float&) { not_provided("cutlas_blockmat_batched", "-M 12"); }
void cublas_dense_multiplyAB(int, int, DataT*, DataT*, int, DataT_C*, float&) { not_provided("cublas_dense_multiplyAB", "-M 1"); }
int cutlass_dense_multiplyAB(int, int, DataT*, int, DataT*, float, float, DataT_C*, float&) { not_provided("cutlass_dense_multiplyAB", "-M 9"); }
DataT* csr2dn(CSR&) { not_provided("csr2dn", "-M 1 / -M 9"); }
