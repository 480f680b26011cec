// TEST HARNESS (built by integration/Makefile where the reference checkout exists, run by
// tests/test_integration_gpu.py on the GPU box).
//
// Calls the reference-facing signatures of integration/cuda_utilities_b200.cpp exactly like
// test/cuda/cuda_multiply.cpp does -- same structs, same leading dimensions, same buffers -- and
// compares every C with the reference's own CPU multiplies IN PROCESS:
//   cublas_fixed_blocks_multiply / cublas_blockmat_batched   vs VBR::multiply (src/general/vbr.cpp:323-372)
//   cublas_blockmat_multiplyBA                               vs a dense product built from the VBR arrays
//   bellpack_blockmat_multiplyAB (Blocked-ELL, row-major)    vs VBR::multiply
//   cusparse_blockmat_multiplyAB (CSR, row-major)            vs CSR::multiply (src/general/csr.cpp:49-65)
// Operands are small integers (pattern-only A, B in {0..3}), exact in fp16 / bf16 / tf32 with fp32
// accumulation, so every comparison is for equality.  This is the numeric check of the shim's
// ld / pointer mapping (ldb = A.cols, ldc = A.rows, BA ld = B_rows, ELL and CSR row-major).
//   usage: shim_check -f matrix.el -P 1 -a 5 -b 16 -B 16 -t 0.6 -c 48 [-F 1]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "blocking.h"
#include "cuda_utilities.h"
#include "cutlass_bellpack_lib.h"
#include "input.h"
#include "matrices.h"

static int report(const char* what, const std::vector<float>& got, const std::vector<float>& want) {
  double worst = 0;
  for (size_t i = 0; i < want.size(); ++i) {
    const double d = got[i] > want[i] ? got[i] - want[i] : want[i] - got[i];
    if (d > worst) worst = d;
  }
  printf("shim_check %-32s max_abs_diff %g over %zu entries\n", what, worst, want.size());
  return worst == 0 ? 0 : 1;
}

int main(int argc, char* argv[]) {
  CLineReader cli(argc, argv);
  CSR cmat(cli);
  BlockingEngine engine(cli);
  engine.GetGrouping(cmat);
  VBR v;
  v.fill_from_CSR_inplace(cmat, engine.grouping_result, cli.col_block_size_, cli.row_block_size_, cli.force_fixed_size);
  const int n = cli.B_cols_;
  const long rows = v.rows, cols = v.cols;
  int bad = 0;
  float dt = 0;

  // B as the VBR paths read it: column-major, ld = cols (+ block_col_size of slack: VBR::multiply reads
  // up to w-1 floats past the last column when cols % w != 0, vbr.cpp:362)
  std::vector<float> Bcm(static_cast<size_t>(cols) * n + v.block_col_size, 0.f), Brm(static_cast<size_t>(cols) * n);
  for (long k = 0; k < cols; ++k)
    for (int j = 0; j < n; ++j) {
      const float x = static_cast<float>((k * 7 + j * 3) % 4);
      Bcm[k + static_cast<size_t>(j) * cols] = x;
      Brm[static_cast<size_t>(k) * n + j] = x;
    }
  std::vector<float> Cref(static_cast<size_t>(rows) * n, 0.f), C(Cref.size(), 0.f);
  v.multiply(Bcm.data(), n, Cref.data());

  cublas_fixed_blocks_multiply(v, Bcm.data(), n, C.data(), dt, 4);                       // -M 4
  bad += report("cublas_fixed_blocks_multiply", C, Cref);
  std::fill(C.begin(), C.end(), 0.f);
  cublas_blockmat_batched(v, Bcm.data(), n, C.data(), dt);                               // -M 7
  bad += report("cublas_blockmat_batched", C, Cref);

  // -M 6: C (n x cols) = B (n x rows) * A, both column-major with ld = n; A rows in blocked order
  {
    std::vector<float> B2(static_cast<size_t>(n) * rows), C2(static_cast<size_t>(n) * cols, 0.f), want(C2.size(), 0.f);
    for (size_t i = 0; i < B2.size(); ++i) B2[i] = static_cast<float>((i * 5 + 1) % 4);
    const float* blk = v.mab;
    const intT* jab = v.jab;
    for (intT ib = 0; ib < v.block_rows; ++ib) {
      const intT h = v.row_part[ib + 1] - v.row_part[ib];
      for (intT q = 0; q < v.nzcount[ib]; ++q, blk += h * v.block_col_size) {
        const intT jb = *jab++;
        for (intT c = 0; c < v.block_col_size && jb * v.block_col_size + c < cols; ++c)
          for (intT r = 0; r < h; ++r) {
            const float a = blk[r + c * h];
            if (a == 0.f) continue;
            for (int i = 0; i < n; ++i)
              want[i + static_cast<size_t>(n) * (jb * v.block_col_size + c)] += B2[i + static_cast<size_t>(n) * (v.row_part[ib] + r)] * a;
          }
      }
    }
    cublas_blockmat_multiplyBA(v, B2.data(), n, C2.data(), dt, 4);
    bad += report("cublas_blockmat_multiplyBA", C2, want);
  }

  // -M 3 / -M 8: Blocked-ELL needs square fixed blocks on padded dimensions (cuda_utilities.cpp:1664-1670)
  if (cli.force_fixed_size && cli.col_block_size_ == cli.row_block_size_ && rows % cli.col_block_size_ == 0 &&
      cols % cli.col_block_size_ == 0) {
    std::vector<float> Crm(static_cast<size_t>(rows) * n, 0.f), want(Crm.size());
    for (long r = 0; r < rows; ++r)
      for (int j = 0; j < n; ++j) want[static_cast<size_t>(r) * n + j] = Cref[r + static_cast<size_t>(j) * rows];
    bellpack_blockmat_multiplyAB(&v, Brm.data(), n, Crm.data(), n, dt, 0);
    bad += report("bellpack_blockmat_multiplyAB", Crm, want);
    std::fill(Crm.begin(), Crm.end(), 0.f);
    bellpack_cutlass_multiplyAB(&v, Brm.data(), n, Crm.data(), n, dt, 0);
    bad += report("bellpack_cutlass_multiplyAB", Crm, want);
  } else {
    printf("shim_check Blocked-ELL paths skipped (need -F 1 and -b == -B)\n");
  }

  // -M 2: CSR x dense, B and C row-major; CSR::multiply indexes B with `rows` as leading dimension
  // (csr.cpp:61), so it is the reference only for square A
  if (cmat.rows == cmat.cols) {
    const long m = cmat.rows;
    std::vector<float> Bsq_cm(static_cast<size_t>(m) * n), Bsq_rm(Bsq_cm.size()), Ccsr(Bsq_cm.size(), 0.f), got(Bsq_cm.size(), 0.f),
        want(Bsq_cm.size());
    for (long k = 0; k < m; ++k)
      for (int j = 0; j < n; ++j) {
        const float x = static_cast<float>((k * 3 + j) % 4);
        Bsq_cm[k + static_cast<size_t>(j) * m] = x;
        Bsq_rm[static_cast<size_t>(k) * n + j] = x;
      }
    cmat.multiply(Bsq_cm.data(), n, Ccsr.data());
    for (long r = 0; r < m; ++r)
      for (int j = 0; j < n; ++j) want[static_cast<size_t>(r) * n + j] = Ccsr[r + static_cast<size_t>(j) * m];
    cusparse_blockmat_multiplyAB(cmat, Bsq_rm.data(), n, got.data(), n, dt);
    bad += report("cusparse_blockmat_multiplyAB", got, want);
  } else {
    printf("shim_check CSR path skipped (CSR::multiply needs a square matrix)\n");
  }
  printf("shim_check %s\n", bad ? "FAILED" : "OK");
  return bad ? 1 : 0;
}
