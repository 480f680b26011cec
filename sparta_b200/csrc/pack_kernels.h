#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sched_types.h"

namespace sparta {

// image_bytes: total bytes of the images the jobs write (0 if unknown)
cudaError_t pack_a_images(const float* src_dev, const PackJob* jobs_dev, int64_t n_jobs,
                          uint8_t* dst_dev, int precision, cudaStream_t stream, int64_t image_bytes = 0);

cudaError_t convert_b(const float* src_dev, int64_t ld_src, int row_major, void* dst_dev,
                      int64_t ldk, int64_t k_total, int64_t n, int precision,
                      cudaStream_t stream, bool keep_fp32 = false);

// dst[idx[i]] = val[i]: rebuilds a mostly-zero fp32 source from its nonzeros (dst zeroed by the caller)
cudaError_t scatter_values(const int64_t* idx_dev, const float* val_dev, int64_t nnz, float* dst_dev,
                           cudaStream_t stream);

// val[i] = the value of val[i] rounded to the operand precision (kept as fp32)
cudaError_t round_values(float* val_dev, int64_t n, int precision, cudaStream_t stream);

// dst[row_map[r]] = src row r for r < rows: C back in the ORIGINAL row order (row_map = the
// reference's get_permutation, utilities.cpp:8-20).  Both matrices have n columns and arbitrary
// element strides (row stride sr, column stride sj).
cudaError_t permute_rows(const float* src, int64_t src_sr, int64_t src_sj, float* dst, int64_t dst_sr,
                         int64_t dst_sj, const int64_t* row_map_dev, int64_t rows, int64_t n,
                         cudaStream_t stream);

}  // namespace sparta
