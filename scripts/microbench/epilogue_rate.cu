// Microbenchmark: what bounds the drain of a 512-column fp32 accumulator (TMEM -> registers -> C)?
// One CTA per SM, 4 drain warps (one per TMEM lane quarter) like the SpMM epilogue; each variant
// drains 512 columns `reps` times and reports SM cycles per drain (median over CTAs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epilogue_rate epilogue_rate.cu && ./epilogue_rate
// C is column-major [128 columns per CTA][ldc rows]; rows_per_seg rows are contiguous per column.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void st16_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// variant bits: 1 zero TMEM behind the load, 2 direct 16-byte stores (thread = column),
//               4 transposed stores through smem (4 lanes = 64 bytes of one column), 8 pipelined loads,
//               16 x32 loads (32 columns per step), 32 stores with .cs (streaming) hint,
//               64 every repetition writes rows it has not written before (cold lines, like the
//               real epilogue), 128 warps 0 and 1 spin on an mbarrier like the idle producer / MMA warps,
//               256 warp 0 streams 32 KB bulk copies from global memory into shared memory for the
//               whole drain (every SM pulling from L2 like the SpMM kernel's producers),
//               1024 TMA stores: 32 x 32 boxes staged in swizzled shared memory, 2 tiles per warp
__global__ void __launch_bounds__(192, 1) drain(const __grid_constant__ CUtensorMap tmap_c, int variant, int reps,
                                                float* C, long long ldc, int seg_rows, long long* out,
                                                const uint8_t* gsrc, long long gsrc_bytes) {
  extern __shared__ __align__(1024) uint8_t dyn[];
  __shared__ uint64_t copy_bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t never_bar;
  __shared__ volatile int stop;
  __shared__ __align__(16) float stage[4][32 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&never_bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&copy_bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_ptr;
  long long t0 = 0, t1 = 0;
  if (warp >= 2) {
    const int q = warp & 3;
    const uint32_t t_lane = tbase + ((uint32_t)(q * 32) << 16);
    for (int c = 0; c < 512; c += 16) st16_zero(t_lane + c);
    wait_st();
    float* stg = stage[warp - 2];
    const int j = q * 32 + lane;                       // column of this CTA's 128
    float* Ccta = C + (long long)blockIdx.x * 128 * ldc;
    // rows: segment s (seg_rows contiguous rows) lives at row offset s * 1024 (scattered like block-rows)
    int rep_now = 0;
    auto row_of = [&](int col) {
      return (long long)(col / seg_rows) * 1024 + (col % seg_rows) + ((variant & 64) ? (long long)rep_now * seg_rows : 0);
    };
    auto store16 = [&](const uint32_t* v, int col) {
      if (variant & 2) {
        float4* d = reinterpret_cast<float4*>(Ccta + (long long)j * ldc + row_of(col));
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 o = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
          if (variant & 32) __stcs(d + g, o); else d[g] = o;
        }
      } else if (variant & 4) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<float4*>(stg + lane * 16 + 4 * (g ^ ((lane >> 1) & 3))) =
              make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
        __syncwarp();
        const int gl = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int jj = 8 * i + (lane >> 2);
          const float4 o = *reinterpret_cast<const float4*>(stg + jj * 16 + 4 * (gl ^ ((jj >> 1) & 3)));
          float4* d = reinterpret_cast<float4*>(Ccta + (long long)(q * 32 + jj) * ldc + row_of(col) + 4 * gl);
          if (variant & 32) __stcs(d, o); else *d = o;
        }
        __syncwarp();
      } else {
        // no stores: keep the values alive
        uint32_t x = 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) x ^= v[r];
        if (x == 0x12345678u) Ccta[0] = 1.0f;
      }
    };
    __syncwarp();
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      rep_now = rep;
      if (variant & 1024) {
        // staging tiles: after the 128 KB the bulk-copy stream uses; 2 x 4 KB per warp, 1024-aligned
        const uint32_t dyn_base = (smem_u32(dyn) + 1023u) & ~1023u;
        const uint32_t tiles = dyn_base + 131072u + (warp - 2) * 8192u;
        for (int c = 0; c < 512; c += 32) {
          const uint32_t tile = tiles + ((c >> 5) & 1) * 4096u;
          uint32_t v[32];
          ld32(t_lane + c, v);
          // the tile used two steps ago must have been read by the copy engine
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          wait_ld();
          if (variant & 1) { st16_zero(t_lane + c); st16_zero(t_lane + c + 16); }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + lane * 128u + ((k ^ (lane & 7)) << 4)),
                         "r"(v[4 * k]), "r"(v[4 * k + 1]), "r"(v[4 * k + 2]), "r"(v[4 * k + 3]) : "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            const int row = (int)row_of(c);
            const int col = blockIdx.x * 128 + q * 32;
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(&tmap_c)),
                         "r"(row), "r"(col), "r"(tile) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
      } else if (variant & 16) {
        for (int c = 0; c < 512; c += 32) {
          uint32_t v[32];
          ld32(t_lane + c, v);
          wait_ld();
          if (variant & 1) { st16_zero(t_lane + c); st16_zero(t_lane + c + 16); }
          store16(v, c);
          store16(v + 16, c + 16);
        }
      } else if (variant & 8) {
        uint32_t va[16], vb[16];
        ld16(t_lane, va);
        for (int c = 0; c < 512; c += 32) {
          wait_ld();
          ld16(t_lane + c + 16, vb);
          if (variant & 1) st16_zero(t_lane + c);
          store16(va, c);
          wait_ld();
          if (c + 32 < 512) ld16(t_lane + c + 32, va);
          if (variant & 1) st16_zero(t_lane + c + 16);
          store16(vb, c + 16);
        }
      } else {
        for (int c = 0; c < 512; c += 16) {
          uint32_t v[16];
          ld16(t_lane + c, v);
          wait_ld();
          if (variant & 1) st16_zero(t_lane + c);
          store16(v, c);
        }
      }
      wait_st();
    }
    t1 = clock64();
    if (warp == 2 && lane == 0) out[blockIdx.x] = (t1 - t0) / reps;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 2 && lane == 0) stop = 1;
  } else if ((variant & 256) && warp == 0) {
    // 4 x 32 KB copies in flight, each from a different place of a buffer much larger than L2's share
    uint32_t phase = 0;
    unsigned long long pos = (unsigned long long)blockIdx.x * 7919ull * 32768ull;
    while (!stop) {
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&copy_bar)), "r"(4u * 32768u) : "memory");
        for (int k = 0; k < 4; ++k) {
          pos = (pos + 32768ull * 613ull) % (unsigned long long)(gsrc_bytes - 32768);
          pos &= ~127ull;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dyn) + k * 32768u), "l"(gsrc + pos), "r"(32768u), "r"(smem_u32(&copy_bar)) : "memory");
        }
      }
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&copy_bar)), "r"(phase) : "memory");
      phase ^= 1u;
      __syncwarp();
    }
  } else if (variant & 128) {
    // like the producer / MMA warps of the SpMM kernel while an accumulator drains
    while (!stop) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&never_bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

int main() {
  const long long ldc = 65536;
  const int grid = 148;
  float* C;
  long long* out;
  cudaMalloc(&C, sizeof(float) * ldc * 128 * grid);
  cudaMalloc(&out, sizeof(long long) * grid);
  uint8_t* gsrc;
  const long long gsrc_bytes = 1ll << 30;
  cudaMalloc(&gsrc, gsrc_bytes);
  cudaMemset(gsrc, 0, gsrc_bytes);
  cudaFuncSetAttribute(drain, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  CUtensorMap tmap;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)ldc, (cuuint64_t)128 * grid};
    const cuuint64_t strides[1] = {(cuuint64_t)ldc * 4};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeFn)sym)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map encode failed %d\n", (int)r); return 1; }
  }
  struct V { int bits; const char* name; };
  const V vs[] = {
      {0, "tcgen05.ld x16 + wait only"},
      {1, "ld x16 + tcgen05.st zero"},
      {16, "ld x32 + wait only"},
      {2, "ld x16 + direct 16 B stores (thread = column)"},
      {1 | 2, "ld x16 + zero + direct stores (first kernel)"},
      {1 | 4, "ld x16 + zero + transposed stores"},
      {1 | 4 | 8, "pipelined ld + zero + transposed stores (current kernel)"},
      {4 | 8, "pipelined ld + transposed stores, no zeroing"},
      {1 | 4 | 8 | 32, "current + st.cs"},
      {1 | 4 | 16, "ld x32 + zero + transposed stores"},
      {1 | 2 | 64, "first kernel, cold lines"},
      {1 | 4 | 8 | 64, "current kernel, cold lines"},
      {1 | 4 | 8 | 64 | 128, "current kernel, cold lines, spinning warps 0/1"},
      {1 | 4 | 8 | 128, "current kernel, warm lines, spinning warps 0/1"},
      {1 | 4 | 8 | 64 | 512, "current kernel, cold lines, 200 KB of shared memory (small L1)"},
      {1 | 2 | 64 | 512, "first kernel, cold lines, 200 KB of shared memory"},
      {1 | 4 | 8 | 64 | 256 | 512, "current kernel, cold lines, every SM streaming bulk copies"},
      {1 | 2 | 64 | 256 | 512, "first kernel, cold lines, every SM streaming bulk copies"},
      {256 | 512, "TMEM ld only, every SM streaming bulk copies"},
      {1 | 64 | 512 | 1024, "TMA stores (32x32 boxes), cold lines"},
      {1 | 64 | 256 | 512 | 1024, "TMA stores (32x32 boxes), cold lines, every SM streaming bulk copies"},
  };
  for (int active : {148}) {
    for (const V& v : vs) {
      for (int seg_rows : {64}) {
        const size_t dsm = (v.bits & 512) ? 200 * 1024 : 0;
        drain<<<active, 192, dsm>>>(tmap, v.bits, 4, C, ldc, seg_rows, out, gsrc, gsrc_bytes);
        cudaDeviceSynchronize();
        drain<<<active, 192, dsm>>>(tmap, v.bits, 16, C, ldc, seg_rows, out, gsrc, gsrc_bytes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(active);
        cudaMemcpy(h.data(), out, sizeof(long long) * active, cudaMemcpyDeviceToHost);
        std::sort(h.begin(), h.end());
        printf("CTAs %3d  %-58s %7lld cycles / 512 columns (min %lld max %lld)  %.1f B/clk/SM\n", active, v.name,
               h[active / 2], h[0], h[active - 1], 262144.0 / h[active / 2]);
      }
    }
  }
  return 0;
}
