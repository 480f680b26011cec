// Microbenchmark: issue rate / execution time of tcgen05.mma (kind::f16, bf16 in, fp32 acc) from
// shared-memory operands in the K-major SWIZZLE_128B layout the SpMM kernel uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

// mode 0: same operands every MMA; mode 1: walk K (+32 B) and alternate accumulator halves like the SpMM loop
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// warps: 0 MMA issuer, 1 copy engine driver (if copy), 2..5 pollers (if poll_lanes > 0)
__global__ void __launch_bounds__(192, 1) bench(int n_mma, int N, int mode, int b_rows_total, long long* out,
                                                 int poll_lanes, int copy, const uint8_t* gsrc, int copy_bytes) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint64_t never_bar;
  __shared__ uint64_t copy_bar;
  __shared__ volatile int stop;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 512 * 128) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3f803f80u;  // bf16 1.0 pairs
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&never_bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&copy_bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (warp == 0) {
    const uint32_t a_addr = base;                 // 128 rows x 128 B panel
    const uint32_t b_addr = base + 16384;         // b_rows_total rows x 128 B
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
    const int n_runs = b_rows_total / N;
    long long t0 = clock64();
    int issued = 0;
    while (issued < n_mma) {
      for (int r = 0; r < n_runs && issued < n_mma; ++r) {
        const uint64_t ad = smem_desc(a_addr);
        const uint64_t bd = smem_desc(b_addr + (mode ? r * N * 128 : 0));
        const uint32_t d = tmem + (mode ? (r * N) % 512 : 0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) mma(d, ad + (mode ? 2 * k : 0), bd + (mode ? 2 * k : 0), idesc);
        }
        __syncwarp();
        issued += 4;
      }
    }
    long long t1 = clock64();
    if (elect_one()) commit(smem_u32(&bar));
    __syncwarp();
    while (!try_wait(smem_u32(&bar), 0)) {}
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; stop = 1; }
  } else if (warp == 1) {
    if (copy && (threadIdx.x & 31) == 0) {
      // stream copy_bytes chunks from global into a scratch smem region, back to back
      const uint32_t dst = base + 16384 + 512 * 128;
      uint32_t phase = 0;
      long long off = 0;
      while (!stop) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&copy_bar)), "r"((uint32_t)copy_bytes) : "memory");
        bulk_load(dst, gsrc + (size_t)blockIdx.x * (1 << 20) + off, copy_bytes, smem_u32(&copy_bar));
        while (!try_wait(smem_u32(&copy_bar), phase)) {}
        phase ^= 1;
        off = (off + copy_bytes) & ((1 << 20) - 1);
      }
    }
  } else {
    if ((int)(threadIdx.x & 31) < poll_lanes) {
      while (!stop) { if (try_wait(smem_u32(&never_bar), 0)) break; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 148 * 2 * sizeof(long long));
  long long h[296];
  const int smem = 1024 + 16384 + 512 * 128 + 65536 + 1024;
  uint8_t* gsrc; cudaMalloc(&gsrc, 148ull << 20); cudaMemset(gsrc, 0, 148ull << 20);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 8192;
  int grids[] = {1, 148};
  int Ns[] = {64, 128, 256};
  struct Cfg { int poll_lanes, copy, copy_bytes; const char* name; };
  Cfg cfgs[] = {{0, 0, 0, "quiet"}, {1, 0, 0, "4 warps x 1 lane polling"}, {32, 0, 0, "4 warps x 32 lanes polling"},
                {0, 1, 32768, "32KB bulk copies"}, {0, 1, 49152, "48KB bulk copies"}, {32, 1, 49152, "poll32 + 48KB copies"}};
  for (Cfg c : cfgs) { printf("--- %s\n", c.name);
  for (int g : grids) for (int mode = 1; mode < 2; ++mode) for (int N : Ns) {
    bench<<<g, 192, smem>>>(n_mma, N, mode, 512, d_out, c.poll_lanes, c.copy, gsrc, c.copy_bytes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d_out, g * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    double issue = 0, total = 0;
    for (int i = 0; i < g; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
    issue /= g; total /= g;
    printf("grid %3d mode %d M=128 N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (ideal %.0f) -> %.0f MAC/cyc/SM\n", g, mode, N,
           issue / n_mma, total / n_mma, N / 2.0, 128.0 * N * 16 * n_mma / total);
  } }
  return 0;
}
