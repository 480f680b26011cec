// CSR x dense SpMM for sm_100a -- the `-M 2` route of the reference
// (cusparse_blockmat_multiplyAB -> cusparse_gemm_custom -> cusparseSpMM(CSR_ALG2),
// src/cuda/cuda_utilities.cpp:1251-1431, :1479-1493).
//
// This product has no dense blocks to feed a tensor core: per nonzero it reads one row of B
// (n contiguous elements) and adds it, scaled, into one row of C.  It is a gather bound by the
// L2 -> SM path, so the kernel is organised around bytes, not MMAs:
//
//   * B is kept row-major [cols][ldn] in the compute precision (bf16/fp16: 2 bytes per element;
//     the "tf32" precision selects plain fp32 rows and IEEE fp32 multiply-then-add in the
//     reference's own summation order, which makes that mode bit-identical to CSR::multiply,
//     src/general/csr.cpp:49-65).
//   * one warp owns one (row, 256-column tile): a lane holds 8 consecutive columns, loads them
//     with one (2-byte types) or two (fp32) 16-byte requests per nonzero and keeps 8 fp32
//     accumulators in registers; column indices and values are fetched 32 at a time, coalesced,
//     and broadcast by shuffle; the nonzero loop is unrolled so 8 requests per lane are in flight.
//   * rows longer than kCsrHeavyNnz entries get a whole CTA (8 warps on 8 slices of the row).
//   * rows are visited in descending-nnz order (row_order, built on the host) so the 8 warps of
//     a CTA carry similar work and the heavy rows start first; blockIdx.y walks the column tiles
//     outermost, so the B slab of one tile (cols x 256 x 2 bytes = 33 MB at cols = 65536) stays
//     L2-resident while every row streams past it.
//   * C is written once with 16-byte stores (row-major) -- no atomics, no read-modify-write
//     unless accumulate is requested.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "csr_kernel.h"

namespace sparta {

// One nonzero's 8 columns as they sit in memory: one 16-byte word for 2-byte types, two for fp32.
template <int kPrec>
struct Raw8 {
  uint4 lo, hi;
};
template <int kPrec>
__device__ __forceinline__ Raw8<kPrec> load8(const void* base, int64_t elem) {
  Raw8<kPrec> r;
  if constexpr (kPrec == PREC_TF32) {
    const uint4* p = reinterpret_cast<const uint4*>(static_cast<const float*>(base) + elem);
    r.lo = __ldg(p);
    r.hi = __ldg(p + 1);
  } else {
    r.lo = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + elem));
    r.hi = make_uint4(0, 0, 0, 0);
  }
  return r;
}

template <int kPrec>
__device__ __forceinline__ void axpy8(float a, const Raw8<kPrec>& r, float (&acc)[8]) {
  float b[8];
  if constexpr (kPrec == PREC_TF32) {
    b[0] = __uint_as_float(r.lo.x); b[1] = __uint_as_float(r.lo.y);
    b[2] = __uint_as_float(r.lo.z); b[3] = __uint_as_float(r.lo.w);
    b[4] = __uint_as_float(r.hi.x); b[5] = __uint_as_float(r.hi.y);
    b[6] = __uint_as_float(r.hi.z); b[7] = __uint_as_float(r.hi.w);
  } else {
    const uint32_t w[4] = {r.lo.x, r.lo.y, r.lo.z, r.lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (kPrec == PREC_BF16) {   // bf16 -> fp32 is a 16-bit shift
        b[2 * i] = __uint_as_float(w[i] << 16);
        b[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        b[2 * i] = f.x;
        b[2 * i + 1] = f.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if constexpr (kPrec == PREC_TF32)
      acc[i] = __fadd_rn(acc[i], __fmul_rn(a, b[i]));   // the reference's own rounding sequence
    else
      acc[i] = fmaf(a, b[i], acc[i]);
  }
}

// acc += sum over the nonzeros [beg, end) of one row, in ascending order, for this lane's 8 columns.
// The whole warp takes part (shuffles); lanes past the last stored column pass live = false.
template <int kPrec>
__device__ __forceinline__ void row_slice(const CsrParams& p, int64_t beg, int64_t end, int j0,
                                          bool live, int lane, float (&acc)[8]) {
  constexpr int kUnroll = (kPrec == PREC_TF32) ? 4 : 8;   // 8 x 16-byte requests in flight per lane
  for (int64_t q0 = beg; q0 < end; q0 += 32) {
    const int cnt = static_cast<int>(min(static_cast<int64_t>(32), end - q0));
    int32_t my_c = 0;
    float my_a = 0.f;
    if (lane < cnt) {
      my_c = __ldg(p.colind + q0 + lane);
      my_a = __ldg(p.val + q0 + lane);
    }
    int i = 0;
    for (; i + kUnroll <= cnt; i += kUnroll) {
      Raw8<kPrec> b[kUnroll];
      float a[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int32_t c = __shfl_sync(0xFFFFFFFFu, my_c, i + u);
        a[u] = __shfl_sync(0xFFFFFFFFu, my_a, i + u);
        if (live) b[u] = load8<kPrec>(p.B, static_cast<int64_t>(c) * p.ldn + j0);
      }
      if (live) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) axpy8<kPrec>(a[u], b[u], acc);
      }
    }
    for (; i < cnt; ++i) {
      const int32_t c = __shfl_sync(0xFFFFFFFFu, my_c, i);
      const float a = __shfl_sync(0xFFFFFFFFu, my_a, i);
      if (live) axpy8<kPrec>(a, load8<kPrec>(p.B, static_cast<int64_t>(c) * p.ldn + j0), acc);
    }
  }
}

__device__ __forceinline__ void store8(const CsrParams& p, int32_t row, int j0, const float (&acc)[8]) {
  float* dst = p.C + static_cast<int64_t>(row) * p.c_sr + static_cast<int64_t>(j0) * p.c_sj;
  if (p.c_sj == 1 && j0 + 8 <= p.n && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    float4 lo = make_float4(acc[0], acc[1], acc[2], acc[3]);
    float4 hi = make_float4(acc[4], acc[5], acc[6], acc[7]);
    float4* d4 = reinterpret_cast<float4*>(dst);
    if (p.accumulate) {
      const float4 a = d4[0], b = d4[1];
      lo.x += a.x; lo.y += a.y; lo.z += a.z; lo.w += a.w;
      hi.x += b.x; hi.y += b.y; hi.z += b.z; hi.w += b.w;
    }
    d4[0] = lo;
    d4[1] = hi;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (j0 + i < p.n) {
        float* d = dst + static_cast<int64_t>(i) * p.c_sj;
        *d = p.accumulate ? (*d + acc[i]) : acc[i];
      }
    }
  }
}

// blockIdx.x < heavy_rows : the CTA owns ONE long row (more than kCsrHeavyNnz entries) of this
//   column tile; its 8 warps take 8 contiguous slices of the row and the partial sums are added
//   in slice order through shared memory (deterministic).  R-MAT rows are heavy-tailed: at
//   BASELINE config #3 the longest row has 14 448 entries and 36 % of all entries sit in rows
//   longer than 512, which a single warp would drag through at a few GB/s.
// otherwise                : the CTA owns 8 short rows, one warp each, summed sequentially in
//   the reference's order.
template <int kPrec>
__global__ void __launch_bounds__(kCsrThreads)
spmm_csr_sm100(const CsrParams p) {
  __shared__ float part[kCsrThreads / 32][kCsrTileJ];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int j0 = blockIdx.y * kCsrTileJ + lane * 8;
  const bool live = j0 < p.ldn;   // ldn is n rounded up to 8: a live lane owns 8 stored columns
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  if (static_cast<int64_t>(blockIdx.x) < p.heavy_rows) {
    const int32_t row = p.row_order[blockIdx.x];
    const int64_t beg = p.rowptr[row], end = p.rowend ? p.rowend[row] : p.rowptr[row + 1];
    const int64_t per = (end - beg + (kCsrThreads / 32) - 1) / (kCsrThreads / 32);
    const int64_t b = min(end, beg + warp * per), e = min(end, b + per);
    row_slice<kPrec>(p, b, e, j0, live, lane, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) part[warp][lane * 8 + i] = acc[i];
    __syncthreads();
    // thread t owns column t of the tile: slices added in ascending order
    const int j = blockIdx.y * kCsrTileJ + threadIdx.x;
    if (j < p.n) {
      float sum = part[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < kCsrThreads / 32; ++w) sum = __fadd_rn(sum, part[w][threadIdx.x]);
      float* d = p.C + static_cast<int64_t>(row) * p.c_sr + static_cast<int64_t>(j) * p.c_sj;
      *d = p.accumulate ? (*d + sum) : sum;
    }
    return;
  }
  const int64_t slot = p.heavy_rows + (static_cast<int64_t>(blockIdx.x) - p.heavy_rows) * (kCsrThreads / 32) + warp;
  if (slot >= p.rows) return;
  const int32_t row = p.row_order[slot];
  row_slice<kPrec>(p, p.rowptr[row], p.rowend ? p.rowend[row] : p.rowptr[row + 1], j0, live, lane, acc);
  if (live) store8(p, row, j0, acc);
}

cudaError_t spmm_csr_launch(const CsrParams& p, int precision, cudaStream_t stream) {
  if (p.rows == 0 || p.n == 0) return cudaSuccess;
  const int64_t gx = p.heavy_rows + (p.rows - p.heavy_rows + (kCsrThreads / 32) - 1) / (kCsrThreads / 32);
  const int64_t gy = (p.ldn + kCsrTileJ - 1) / kCsrTileJ;
  if (gx > 0x7FFFFFFFll || gy > 65535) return cudaErrorInvalidConfiguration;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
  if (precision == PREC_BF16)
    spmm_csr_sm100<PREC_BF16><<<grid, kCsrThreads, 0, stream>>>(p);
  else if (precision == PREC_FP16)
    spmm_csr_sm100<PREC_FP16><<<grid, kCsrThreads, 0, stream>>>(p);
  else
    spmm_csr_sm100<PREC_TF32><<<grid, kCsrThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace sparta
