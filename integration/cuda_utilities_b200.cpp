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

// B column-major (ld = A.cols), C column-major (ld = A.rows), like VBR::multiply (vbr.cpp:331,355)
void vbr_multiply(const VBR& A, DataT* B, int B_cols, DataT_C* C, float& dt, int precision) {
  static_assert(sizeof(intT) == sizeof(int64_t) && sizeof(DataT) == sizeof(float) && sizeof(DataT_C) == sizeof(float),
                "the C ABI takes the reference's default types (definitions.h:4-6)");
  if (sparta_vbr_spmm(A.rows, A.cols, A.block_rows, A.block_col_size,
                      reinterpret_cast<const int64_t*>(A.row_part), reinterpret_cast<const int64_t*>(A.nzcount),
                      reinterpret_cast<const int64_t*>(A.jab), A.mab, B, A.cols, B_cols, C, A.rows, precision, &dt))
    die("sparta_vbr_spmm");
}

// C = B*A: B is B_rows x A.rows, C is B_rows x A.cols, both column-major with ld = B_rows
// (cuda_utilities.cpp:556-559,587-591)
void vbr_multiply_BA(const VBR& A, DataT* B, int B_rows, DataT_C* C, float& dt, int precision) {
  if (sparta_vbr_spmm_BA(A.rows, A.cols, A.block_rows, A.block_col_size,
                         reinterpret_cast<const int64_t*>(A.row_part), reinterpret_cast<const int64_t*>(A.nzcount),
                         reinterpret_cast<const int64_t*>(A.jab), A.mab, B, B_rows, B_rows, C, B_rows, precision, &dt))
    die("sparta_vbr_spmm_BA");
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

// prepare_cusparse_CSR (cuda_utilities.cpp:1433-1477) flattens the reference's array-of-rows CSR
// into rowptr / colind / val; the same flattening here, with the ABI's int64 indices.  The
// reference asks cuSPARSE for CUDA_R_32F compute (:1267), hence the fp32 default.
void cusparse_blockmat_multiplyAB(CSR& A, DataT* B, int B_cols, DataT_C* C, int C_cols, float& dt) {
  std::vector<int64_t> rowptr(static_cast<size_t>(A.rows) + 1, 0);
  for (intT i = 0; i < A.rows; ++i) rowptr[i + 1] = rowptr[i] + A.nzcount[i];
  std::vector<int64_t> colind(static_cast<size_t>(rowptr[A.rows]));
  std::vector<float> val(A.pattern_only ? 0 : colind.size());
  for (intT i = 0; i < A.rows; ++i)
    for (intT q = 0; q < A.nzcount[i]; ++q) {
      colind[rowptr[i] + q] = A.ja[i][q];
      if (!A.pattern_only) val[rowptr[i] + q] = A.ma[i][q];
    }
  if (sparta_csr_spmm(A.rows, A.cols, rowptr.data(), colind.data(), A.pattern_only ? nullptr : val.data(), B, B_cols,
                      B_cols, C, C_cols, precision_or(SPARTA_TF32), &dt))
    die("sparta_csr_spmm");
}

// ---- debug printers the CLI calls at -v 3 ------------------------------------------------------

void pico_print_DnM(const char* Cname, int Cn, int Cm, DataT_C* C) {
  printf("Dense matrix %s (%d x %d):\n", Cname, Cn, Cm);
  for (int i = 0; i < Cn; ++i) {
    for (int j = 0; j < Cm; ++j) printf("%g ", static_cast<double>(C[i * Cm + j]));
    printf("\n");
  }
}

// ---- outside the hot path ----------------------------------------------------------------------

void cutlas_blockmat_batched(const VBR&, DataT*, int, DataT_C*, float&) { not_provided("cutlas_blockmat_batched", "-M 12"); }
void cublas_dense_multiplyAB(int, int, DataT*, DataT*, int, DataT_C*, float&) { not_provided("cublas_dense_multiplyAB", "-M 1"); }
int cutlass_dense_multiplyAB(int, int, DataT*, int, DataT*, float, float, DataT_C*, float&) { not_provided("cutlass_dense_multiplyAB", "-M 9"); }
DataT* csr2dn(CSR&) { not_provided("csr2dn", "-M 1 / -M 9"); }
