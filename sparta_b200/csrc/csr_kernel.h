// Launch interface of the sm_100a CSR x dense kernel (csr_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sched_types.h"

namespace sparta {

struct CsrParams {
  const int64_t* rowptr;     // [rows + 1]: row r owns entries [rowptr[r], rowptr[r + 1]) ...
  const int64_t* rowend;     // ... or [rowptr[r], rowend[r]) when rowend is given (a column range of the rows)
  const int32_t* colind;     // [nnz], ascending inside a row like the reference's CSR
  const float*   val;        // [nnz] (ones for pattern-only matrices, csr.cpp:59)
  const int32_t* row_order;  // rows in descending-nnz order (host, stable)
  const void*    B;          // row-major [cols][ldn] in the compute precision (fp32 for PREC_TF32)
  float*         C;
  int64_t        c_sr;       // element stride of C between rows
  int64_t        c_sj;       // element stride of C between columns
  int64_t        rows;
  int64_t        heavy_rows; // the first heavy_rows entries of row_order have more than kCsrHeavyNnz entries
  int32_t        n;          // columns of B and C
  int32_t        ldn;        // n rounded up to 8 (zero padded)
  int32_t        accumulate; // 1: C += A*B, 0: C = A*B
};

constexpr int kCsrThreads = 256;   // 8 warps = 8 rows of one column tile
constexpr int kCsrTileJ   = 256;   // columns per warp pass (8 per lane)
constexpr int kCsrHeavyNnz = 512;  // rows up to this length are summed sequentially by one warp

cudaError_t spmm_csr_launch(const CsrParams& p, int precision, cudaStream_t stream);

}  // namespace sparta
