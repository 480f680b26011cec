#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sched_types.h"

namespace sparta {

cudaError_t pack_a_images(const float* src_dev, const PackJob* jobs_dev, int64_t n_jobs,
                          uint8_t* dst_dev, int precision, cudaStream_t stream);

cudaError_t convert_b(const float* src_dev, int64_t ld_src, int row_major, void* dst_dev,
                      int64_t ldk, int64_t k_total, int64_t n, int precision,
                      cudaStream_t stream, bool keep_fp32 = false);

}  // namespace sparta
