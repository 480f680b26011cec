// TEST INFRASTRUCTURE ONLY -- never linked into libsparta_b200.
//
// Thin extern "C" driver around the UNMODIFIED reference sources
// (/root/reference/src/general/*.cpp, compiled where they lie by oracle/Makefile
// into oracle/_ref/libsparta_ref.so).  It runs the reference's own front half
// exactly as test/general/TEST_blocking_VBR.cpp:10-41 does
//   CLineReader -> CSR(cli) -> BlockingEngine::GetGrouping -> VBR::fill_from_CSR_inplace
// and hands the resulting arrays out so the tests can pin the restatement
// (oracle/sparta_oracle.cpp) and the product host layer against the real thing.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "blocking.h"
#include "definitions.h"
#include "input.h"
#include "matrices.h"
#include "utilities.h"

extern "C" {

struct RefResult {
  // CSR after -r reordering
  long csr_rows, csr_cols, csr_nnz;
  long* csr_rowptr;   // [rows+1]
  long* csr_colind;   // [nnz]
  float* csr_val;     // [nnz] (1.0 when pattern-only)
  // blocking
  long* grouping;     // [csr_rows]
  long comparison_counter, merge_counter;
  float average_merge_tau, average_row_distance;
  long VBR_nzcount, VBR_nzblocks_count, VBR_longest_row;
  float VBR_average_height;
  // VBR
  long rows, cols, block_rows, block_cols, block_col_size, nztot, jab_len;
  long* row_part;     // [block_rows+1]
  long* nzcount;      // [block_rows]
  long* jab;          // [jab_len]
  float* mab;         // [nztot]
};

static long* copy_longs(const long* src, size_t n) {
  long* p = static_cast<long*>(malloc((n ? n : 1) * sizeof(long)));
  if (n) memcpy(p, src, n * sizeof(long));
  return p;
}

// argv-style flags exactly as the reference CLIs take them (include/input.h:81).
// fill != 0 also builds the VBR; returns 0 on success.
int ref_run(int argc, char** argv, int fill, RefResult* out) {
  memset(out, 0, sizeof(*out));
  optind = 1;  // CLineReader uses getopt
  try {
    CLineReader cli(argc, argv);
    CSR cmat(cli);
    out->csr_rows = cmat.rows;
    out->csr_cols = cmat.cols;
    out->csr_nnz = cmat.nztot();
    out->csr_rowptr = static_cast<long*>(malloc((cmat.rows + 1) * sizeof(long)));
    out->csr_colind = static_cast<long*>(malloc((out->csr_nnz ? out->csr_nnz : 1) * sizeof(long)));
    out->csr_val = static_cast<float*>(malloc((out->csr_nnz ? out->csr_nnz : 1) * sizeof(float)));
    long pos = 0;
    for (long i = 0; i < cmat.rows; ++i) {
      out->csr_rowptr[i] = pos;
      for (long k = 0; k < cmat.nzcount[i]; ++k) {
        out->csr_colind[pos] = cmat.ja[i][k];
        out->csr_val[pos] = cmat.pattern_only ? 1.0f : cmat.ma[i][k];
        ++pos;
      }
    }
    out->csr_rowptr[cmat.rows] = pos;

    BlockingEngine engine(cli);
    engine.GetGrouping(cmat);
    out->grouping = copy_longs(engine.grouping_result.data(), engine.grouping_result.size());
    out->comparison_counter = engine.comparison_counter;
    out->merge_counter = engine.merge_counter;
    out->average_merge_tau = engine.average_merge_tau;
    out->average_row_distance = engine.average_row_distance;
    engine.CollectBlockingInfo(cmat);
    out->VBR_nzcount = engine.VBR_nzcount;
    out->VBR_nzblocks_count = engine.VBR_nzblocks_count;
    out->VBR_average_height = engine.VBR_average_height;
    out->VBR_longest_row = engine.VBR_longest_row;

    if (fill) {
      VBR v;
      v.fill_from_CSR_inplace(cmat, engine.grouping_result, cli.col_block_size_, cli.row_block_size_,
                              cli.force_fixed_size);
      out->rows = v.rows;
      out->cols = v.cols;
      out->block_rows = v.block_rows;
      out->block_cols = v.block_cols;
      out->block_col_size = v.block_col_size;
      out->nztot = v.nztot;
      long jl = 0;
      for (long b = 0; b < v.block_rows; ++b) jl += v.nzcount[b];
      out->jab_len = jl;
      out->row_part = copy_longs(v.row_part, v.block_rows + 1);
      out->nzcount = copy_longs(v.nzcount, v.block_rows);
      out->jab = copy_longs(v.jab, jl);
      out->mab = static_cast<float*>(malloc((v.nztot ? v.nztot : 1) * sizeof(float)));
      if (v.nztot) memcpy(out->mab, v.mab, v.nztot * sizeof(float));
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_run: %s\n", e.what());
    return 1;
  }
  return 0;
}

void ref_free(RefResult* r) {
  free(r->csr_rowptr); free(r->csr_colind); free(r->csr_val); free(r->grouping);
  free(r->row_part); free(r->nzcount); free(r->jab); free(r->mab);
  memset(r, 0, sizeof(*r));
}

// The reference's serial multiply, VBR::multiply (src/general/vbr.cpp:323-372), on
// caller-provided VBR arrays.  C must be zero-filled by the caller (beta = 1).
void ref_vbr_multiply(long rows, long cols, long block_rows, long block_col_size, long* row_part,
                      long* nzcount, long* jab, float* mab, float* B, int B_cols, float* C) {
  VBR v;
  v.rows = rows;
  v.cols = cols;
  v.block_rows = block_rows;
  v.block_cols = (cols - 1) / block_col_size + 1;
  v.block_col_size = block_col_size;
  v.row_part = row_part;
  v.nzcount = nzcount;
  v.jab = jab;
  v.mab = mab;
  v.nztot = 0;
  v.multiply(B, B_cols, C);
  // the arrays belong to the caller: keep ~VBR from deleting them (VBR::clean, vbr.cpp:13)
  v.rows = 0;
  v.cols = 0;
}

// Distance functions (src/general/blocking.cpp:859, :923) for TEST_similarities-style checks.
float ref_hamming(long* a, long na, long ga, long* b, long nb, long gb, long block_size) {
  std::vector<intT> va(a, a + na);
  return HammingDistanceGroup(va, ga, b, nb, gb, block_size);
}
float ref_jaccard(long* a, long na, long ga, long* b, long nb, long gb, long block_size) {
  std::vector<intT> va(a, a + na);
  return JaccardDistanceGroup(va, ga, b, nb, gb, block_size);
}

}  // extern "C"
