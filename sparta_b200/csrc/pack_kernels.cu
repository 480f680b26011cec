// Upload-side kernels: fp32 host layouts -> the device layouts the SpMM kernel
// consumes.  All are HBM-bound element shuffles (no tensor cores).
//
//  * pack_a_images  : VBR mab blocks (column-major, ld = h, src/general/vbr.cpp:224)
//                     or Blocked-ELL values (row-major, cuda_utilities.cpp:1699-1707)
//                     -> K-major SWIZZLE_128B images in compute precision.
//  * convert_b_*    : B fp32 column-major (ld = cols, vbr.cpp:351) or row-major
//                     (ld = n, cuda_utilities.cpp:1581) -> [n][ldk] k-contiguous.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "pack_kernels.h"

namespace sparta {

constexpr int kPrecRawFp32 = 3;   // convert_b only: copy fp32 unchanged (the CSR kernel's fp32 mode)

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi, int precision) {
  if (precision == PREC_BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// One CTA per job: the right shape when an image is tens of rows (64x64 blocks: 256 items of work).
__global__ void __launch_bounds__(256) pack_a_images_cta_kernel(const float* __restrict__ src,
                                                                const PackJob* __restrict__ jobs,
                                                                uint8_t* __restrict__ dst,
                                                                int precision) {
  const PackJob job = jobs[blockIdx.x];
  const int epc = (precision == PREC_TF32) ? 4 : 8;  // elements per 16-byte chunk
  uint8_t* out = dst + static_cast<size_t>(job.dst_off16) * 16;
  const int total = job.h_pad * 8;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int r = idx % job.h_pad;
    const int c = idx / job.h_pad;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = job.k_lo + c * epc + e;   // block-local k
      v[e] = (e < epc && r < job.h && k >= 0 && k < job.k_w)
                 ? src[job.src_base + static_cast<int64_t>(r) * job.src_rs +
                       static_cast<int64_t>(k) * job.src_ks]
                 : 0.f;
    }
    uint4 o;
    if (precision == PREC_TF32) {
      o = make_uint4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
    } else {
      o = make_uint4(pack2(v[0], v[1], precision), pack2(v[2], v[3], precision),
                     pack2(v[4], v[5], precision), pack2(v[6], v[7], precision));
    }
    const int ro = job.r_base + r;   // row inside the image
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(ro) * 128 + ((c ^ (ro & 7)) << 4)) = o;
  }
}

// One WARP per job (8 jobs per CTA): fused short block-rows produce millions of jobs of a row or
// two, for which a CTA per job spent a second launching blocks.  Lane t walks (row r fastest,
// 16-byte chunk c) so that column-major sources are read coalesced along r.
__global__ void __launch_bounds__(256) pack_a_images_kernel(const float* __restrict__ src,
                                                            const PackJob* __restrict__ jobs,
                                                            int64_t n_jobs,
                                                            uint8_t* __restrict__ dst,
                                                            int precision) {
  const int64_t jid = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (jid >= n_jobs) return;
  const PackJob job = jobs[jid];
  const int epc = (precision == PREC_TF32) ? 4 : 8;  // elements per 16-byte chunk
  uint8_t* out = dst + static_cast<size_t>(job.dst_off16) * 16;
  const int total = job.h_pad * 8;
  for (int idx = threadIdx.x & 31; idx < total; idx += 32) {
    const int r = idx % job.h_pad;
    const int c = idx / job.h_pad;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = job.k_lo + c * epc + e;   // block-local k
      v[e] = (e < epc && r < job.h && k >= 0 && k < job.k_w)
                 ? src[job.src_base + static_cast<int64_t>(r) * job.src_rs +
                       static_cast<int64_t>(k) * job.src_ks]
                 : 0.f;
    }
    uint4 o;
    if (precision == PREC_TF32) {
      o = make_uint4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
    } else {
      o = make_uint4(pack2(v[0], v[1], precision), pack2(v[2], v[3], precision),
                     pack2(v[4], v[5], precision), pack2(v[6], v[7], precision));
    }
    // Swizzle<3,4,3>: 16-byte chunk index XOR (row mod 8) inside each 1024-byte atom
    const int ro = job.r_base + r;   // row inside the image
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(ro) * 128 + ((c ^ (ro & 7)) << 4)) = o;
  }
}

// dst[j][k] = cvt(src[j*ld_src + k])   (source already k-contiguous)
__global__ void convert_b_colmajor_kernel(const float* __restrict__ src, int64_t ld_src,
                                          void* __restrict__ dst, int64_t ldk, int64_t k_total,
                                          int64_t n, int precision) {
  const int64_t j = blockIdx.y;
  const float* s = src + j * ld_src;
  for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < k_total;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = s[k];
    if (precision == kPrecRawFp32)
      reinterpret_cast<float*>(dst)[j * ldk + k] = x;
    else if (precision == PREC_TF32)
      reinterpret_cast<uint32_t*>(dst)[j * ldk + k] = to_tf32(x);
    else if (precision == PREC_BF16)
      reinterpret_cast<__nv_bfloat16*>(dst)[j * ldk + k] = __float2bfloat16_rn(x);
    else
      reinterpret_cast<__half*>(dst)[j * ldk + k] = __float2half_rn(x);
  }
}

// dst[j][k] = cvt(src[k*ld_src + j])   (row-major source: transpose through smem)
__global__ void convert_b_rowmajor_kernel(const float* __restrict__ src, int64_t ld_src,
                                          void* __restrict__ dst, int64_t ldk, int64_t k_total,
                                          int64_t n, int precision) {
  __shared__ float tile[32][33];
  const int64_t k0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int64_t j0 = static_cast<int64_t>(blockIdx.y) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t k = k0 + i, j = j0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < k_total && j < n) ? src[k * ld_src + j] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t j = j0 + i, k = k0 + threadIdx.x;
    if (j < n && k < k_total) {
      const float x = tile[threadIdx.x][i];
      if (precision == kPrecRawFp32)
        reinterpret_cast<float*>(dst)[j * ldk + k] = x;
      else if (precision == PREC_TF32)
        reinterpret_cast<uint32_t*>(dst)[j * ldk + k] = to_tf32(x);
      else if (precision == PREC_BF16)
        reinterpret_cast<__nv_bfloat16*>(dst)[j * ldk + k] = __float2bfloat16_rn(x);
      else
        reinterpret_cast<__half*>(dst)[j * ldk + k] = __float2half_rn(x);
    }
  }
}

cudaError_t pack_a_images(const float* src_dev, const PackJob* jobs_dev, int64_t n_jobs,
                          uint8_t* dst_dev, int precision, cudaStream_t stream, int64_t image_bytes) {
  // mean work per job in 16-byte pieces decides between a CTA and a warp per job
  const bool small_jobs = image_bytes > 0 && image_bytes / 16 / (n_jobs > 0 ? n_jobs : 1) < 128;
  // gridDim.x limit is 2^31-1; stay well below it per launch anyway
  const int64_t kMaxGrid = 1 << 30;
  for (int64_t done = 0; done < n_jobs; done += kMaxGrid) {
    const int64_t g = (n_jobs - done < kMaxGrid) ? (n_jobs - done) : kMaxGrid;
    if (small_jobs)
      pack_a_images_kernel<<<static_cast<unsigned>((g + 7) / 8), 256, 0, stream>>>(src_dev, jobs_dev + done, g,
                                                                                  dst_dev, precision);
    else
      pack_a_images_cta_kernel<<<static_cast<unsigned>(g), 256, 0, stream>>>(src_dev, jobs_dev + done, dst_dev,
                                                                            precision);
  }
  return cudaGetLastError();
}

cudaError_t convert_b(const float* src_dev, int64_t ld_src, int row_major, void* dst_dev,
                      int64_t ldk, int64_t k_total, int64_t n, int precision,
                      cudaStream_t stream, bool keep_fp32) {
  if (k_total == 0 || n == 0) return cudaSuccess;
  const size_t esz = keep_fp32 ? 4 : prec_esize(precision);
  if (keep_fp32) precision = kPrecRawFp32;
  if (!row_major) {
    // gridDim.y <= 65535: loop over column slabs
    for (int64_t j0 = 0; j0 < n; j0 += 65535) {
      const int64_t nj = (n - j0 < 65535) ? (n - j0) : 65535;
      int64_t gx = (k_total + 255) / 256;
      if (gx > 1024) gx = 1024;
      dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(nj));
      convert_b_colmajor_kernel<<<grid, 256, 0, stream>>>(
          src_dev + j0 * ld_src, ld_src, static_cast<uint8_t*>(dst_dev) + j0 * ldk * esz, ldk,
          k_total, nj, precision);
    }
  } else {
    for (int64_t j0 = 0; j0 < n; j0 += 32 * 65535LL) {
      const int64_t nj = (n - j0 < 32 * 65535LL) ? (n - j0) : 32 * 65535LL;
      dim3 grid(static_cast<unsigned>((k_total + 31) / 32), static_cast<unsigned>((nj + 31) / 32));
      convert_b_rowmajor_kernel<<<grid, dim3(32, 8), 0, stream>>>(
          src_dev + j0, ld_src, static_cast<uint8_t*>(dst_dev) + j0 * ldk * esz, ldk, k_total, nj,
          precision);
    }
  }
  return cudaGetLastError();
}

__global__ void scatter_values_kernel(const int64_t* __restrict__ idx, const float* __restrict__ val, int64_t nnz,
                                      float* __restrict__ dst) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += stride)
    dst[idx[i]] = val[i];
}

// grid cap of the grid-stride helper kernels: 16 CTAs per SM of the current device
static int64_t stride_grid_cap() {
  static int64_t cached = 0;
  if (cached) return cached;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    sms = 148;
  cached = static_cast<int64_t>(sms) * 16;
  return cached;
}

cudaError_t scatter_values(const int64_t* idx_dev, const float* val_dev, int64_t nnz, float* dst_dev,
                           cudaStream_t stream) {
  if (nnz <= 0) return cudaSuccess;
  int64_t grid = (nnz + 255) / 256;
  if (grid > stride_grid_cap()) grid = stride_grid_cap();
  scatter_values_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(idx_dev, val_dev, nnz, dst_dev);
  return cudaGetLastError();
}

// The gather rows of a hybrid handle keep their nonzeros in fp32 but ROUNDED to the operand precision,
// so that both kernel families of a handle compute on the same numbers (pack_a_images rounds the
// block images the same way).
__global__ void round_values_kernel(float* __restrict__ val, int64_t n, int precision) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float x = val[i];
    float r;
    if (precision == PREC_TF32) r = __uint_as_float(to_tf32(x));
    else if (precision == PREC_BF16) r = __bfloat162float(__float2bfloat16_rn(x));
    else r = __half2float(__float2half_rn(x));
    val[i] = r;
  }
}

cudaError_t round_values(float* val_dev, int64_t n, int precision, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  int64_t grid = (n + 255) / 256;
  if (grid > stride_grid_cap()) grid = stride_grid_cap();
  round_values_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(val_dev, n, precision);
  return cudaGetLastError();
}

// HBM-bound scatter of C's rows: 2 x rows x n x 4 bytes.  Threads run along whichever dimension is
// contiguous in the source so the reads coalesce; the writes are 4-byte scatters when C is
// column-major (consecutive blocked rows map to non-consecutive original rows) and full lines
// when it is row-major.
__global__ void permute_rows_kernel(const float* __restrict__ src, int64_t src_sr, int64_t src_sj,
                                    float* __restrict__ dst, int64_t dst_sr, int64_t dst_sj,
                                    const int64_t* __restrict__ row_map, int64_t rows, int64_t n) {
  const int64_t total = rows * n;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t r, j;
    if (src_sr == 1) { r = i % rows; j = i / rows; } else { j = i % n; r = i / n; }
    dst[row_map[r] * dst_sr + j * dst_sj] = src[r * src_sr + j * src_sj];
  }
}

cudaError_t permute_rows(const float* src, int64_t src_sr, int64_t src_sj, float* dst, int64_t dst_sr,
                         int64_t dst_sj, const int64_t* row_map_dev, int64_t rows, int64_t n,
                         cudaStream_t stream) {
  if (rows == 0 || n == 0) return cudaSuccess;
  const int64_t total = rows * n;
  int64_t grid = (total + 255) / 256;
  if (grid > stride_grid_cap()) grid = stride_grid_cap();
  permute_rows_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(src, src_sr, src_sj, dst, dst_sr, dst_sj,
                                                                     row_map_dev, rows, n);
  return cudaGetLastError();
}

}  // namespace sparta
