// Microbenchmark: does TMA multicast across a CTA cluster lower the cost of feeding the SAME 16 KB
// B panel to several SMs when the source is L2-resident?  (The SpMM kernel is bound by what the L2
// can hand to the SMs, profiles/r1_spmm_full_summary.md.)
//
// Every CTA runs a 12-stage producer / consumer ring over `iters` panels of 128 rows x 128 bytes
// (the SWIZZLE_128B box the SpMM kernel loads).  Modes:
//   0  unicast, every CTA loads its own panels (no sharing; the L2 -> SM ceiling)
//   1  unicast, the CTAs of a cluster load the SAME panels (sharing left to the L2)
//   2  multicast: each CTA loads 1/c of the panel and multicasts it to all c CTAs of the cluster
//   3  LINEAR bulk copy (cp.async.bulk, no tensor map) of 16 KB contiguous, every CTA its own
//   4  linear bulk copy, multicast: each CTA copies 1/c of the 16 KB to all c CTAs
//   5  per stage one tensor op of 8 KB (64 rows) + one linear bulk op of 8 KB (do the two kinds overlap?)
//   6  per stage two tensor ops of 8 KB;  7  per stage two linear ops of 8 KB
// Prints bytes landed in shared memory per second (aggregate) for cluster sizes 1, 2, 4, 8 and the
// number of co-resident clusters the device grants at the SpMM kernel's shared-memory footprint.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mcast_rate mcast_rate.cu -lcuda && ./mcast_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kStages = 12;
constexpr int kPanel = 16384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  unsigned long long t0 = 0;
  uint32_t spins = 0;
  while (!try_wait(bar, parity)) {
    if (((++spins) & 0xFFF) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (!t0) t0 = t;
      if (t - t0 > 2000000000ull) { printf("mcast_rate: wait timed out cta %d\n", blockIdx.x); __trap(); }
    }
  }
}
__device__ __forceinline__ uint32_t test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void wait_v(uint32_t bar, uint32_t parity, int variant) {
  if (variant == 0) { wait(bar, parity); return; }
  uint32_t spins = 0;
  while (!test_wait(bar, parity)) { if (++spins > (1u << 26)) { printf("mcast_rate: test_wait timed out\n"); __trap(); } }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred != 0;
}

__global__ void __launch_bounds__(64, 1) bench(const __grid_constant__ CUtensorMap tmap_full,
                                               const __grid_constant__ CUtensorMap tmap_slice, int mode, int csize,
                                               int iters, int panels_total, long long* cycles, const uint8_t* src) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[kStages], empty[kStages];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int cluster_id = blockIdx.x / csize;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"((mode == 2 || mode == 4) ? csize : 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const long long t0 = clock64();
  if (warp == 0) {
    uint32_t slot = 0, phase = 0;
    for (int i = 0; i < iters; ++i) {
      if (i >= kStages) wait(smem_u32(&empty[slot]), phase ^ 1u);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&full[slot]);
        const uint32_t dst = base + slot * kPanel;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kPanel) : "memory");
        if (mode >= 5) {
          const int p = ((int)blockIdx.x * iters + i) % panels_total;
          for (int hlf = 0; hlf < 2; ++hlf) {
            const bool tensor = mode == 6 || (mode == 5 && hlf == 0);
            if (tensor)
              asm volatile(
                  "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
                  " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst + hlf * 8192),
                  "l"(reinterpret_cast<uint64_t>(&tmap_slice)), "r"(0), "r"(p * 128 + hlf * 64), "r"(bar)
                  : "memory");
            else
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + hlf * 8192),
                           "l"(src + (size_t)p * kPanel + hlf * 8192), "r"(8192), "r"(bar) : "memory");
          }
        } else if (mode == 3) {
          const int p = ((int)blockIdx.x * iters + i) % panels_total;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                       "l"(src + (size_t)p * kPanel), "r"(kPanel), "r"(bar) : "memory");
        } else if (mode == 4) {
          const int p = (cluster_id * iters + i) % panels_total;
          const uint32_t part = kPanel / csize;
          const uint16_t mask = (uint16_t)((1u << csize) - 1u);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
                       "r"(dst + rank * part), "l"(src + (size_t)p * kPanel + rank * part), "r"(part), "r"(bar), "h"(mask) : "memory");
        } else if (mode == 2) {
          const int p = (cluster_id * iters + i) % panels_total;
          const int rows = 128 / csize;
          const uint16_t mask = (uint16_t)((1u << csize) - 1u);
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
              " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst + rank * rows * 128),
              "l"(reinterpret_cast<uint64_t>(&tmap_slice)), "r"(0), "r"(p * 128 + (int)rank * rows), "r"(bar), "h"(mask)
              : "memory");
        } else {
          const int p = ((mode == 1 ? cluster_id : (int)blockIdx.x) * iters + i) % panels_total;
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
              " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
              "l"(reinterpret_cast<uint64_t>(&tmap_full)), "r"(0), "r"(p * 128), "r"(bar)
              : "memory");
        }
      }
      __syncwarp();
      if (++slot == kStages) { slot = 0; phase ^= 1u; }
    }
  } else {
    uint32_t slot = 0, phase = 0;
    for (int i = 0; i < iters; ++i) {
      wait(smem_u32(&full[slot]), phase);
      if (lane == 0) {
        if (mode == 2 || mode == 4) {
          for (int r = 0; r < csize; ++r) {
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&empty[slot])), "r"(r));
            asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
          }
        } else {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
        }
      }
      __syncwarp();
      if (++slot == kStages) { slot = 0; phase ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// SpMM-like stage: one 16 KB tensor op (the B panel) + `lin_bytes` of linear bulk copy (the A images) in
// pieces of at most 32 KB, 4 stages.
__global__ void __launch_bounds__(64, 1) bench_mix(const __grid_constant__ CUtensorMap tmap_full, int lin_bytes, int tensor_on,
                                                   int iters, int panels_total, long long* cycles, const uint8_t* src, int S, int variant) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[16], empty[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_bytes = 16384 + 32768;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    uint32_t slot = 0, phase = 0;
    for (int i = 0; i < iters; ++i) {
      if (i >= S && variant != 2) wait_v(smem_u32(&empty[slot]), phase ^ 1u, variant);
      if (elect_one()) {
        const uint32_t bar = smem_u32(&full[slot]);
        const uint32_t dst = base + slot * stage_bytes;
        const int p = ((int)blockIdx.x * iters + i) % (panels_total - 4);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((tensor_on ? 16384 : 0) + lin_bytes) : "memory");
        if (tensor_on)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                       "l"(reinterpret_cast<uint64_t>(&tmap_full)), "r"(0), "r"(p * 128), "r"(bar) : "memory");
        if (lin_bytes)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 16384),
                       "l"(src + (size_t)(p + 1) * kPanel), "r"(lin_bytes), "r"(bar) : "memory");
      }
      __syncwarp();
      if (++slot == S) { slot = 0; phase ^= 1u; }
    }
  } else {
    uint32_t slot = 0, phase = 0;
    for (int i = 0; i < iters; ++i) {
      wait_v(smem_u32(&full[slot]), phase, variant == 2 ? 0 : variant);
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
      __syncwarp();
      if (++slot == S) { slot = 0; phase ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

// What serialises the stages?  `per_iter` independent stages (each its own mbarrier, one linear bulk op of
// lin_bytes) are issued back to back per producer iteration by `producers` warps (warp w handles the
// iterations with i % producers == w); ring of S stages.  One consumer warp releases them in order.
__global__ void __launch_bounds__(160, 1) bench_two(int lin_bytes, int per_iter, int producers, int S, int iters,
                                                    int panels_total, long long* cycles, const uint8_t* src) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[16], empty[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  const int groups = iters / per_iter;
  if (warp < producers) {
    for (int g = warp; g < groups; g += producers) {
      for (int u = 0; u < per_iter; ++u) {
        const int i = g * per_iter + u;
        const uint32_t slot = i % S, phase = (i / S) & 1;
        if (i >= S) wait(smem_u32(&empty[slot]), phase ^ 1u);
        if (elect_one()) {
          const uint32_t bar = smem_u32(&full[slot]);
          const int p = ((int)blockIdx.x * iters + i) % (panels_total - 4);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(lin_bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + slot * 16384),
                       "l"(src + (size_t)p * kPanel), "r"(lin_bytes), "r"(bar) : "memory");
        }
        __syncwarp();
      }
    }
  } else if (warp == 4) {
    for (int i = 0; i < groups * per_iter; ++i) {
      const uint32_t slot = i % S, phase = (i / S) & 1;
      wait(smem_u32(&full[slot]), phase);
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
      __syncwarp();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  EncodeFn encode = (EncodeFn)sym;
  const int panels_total = argc > 1 ? atoi(argv[1]) : 4096;   // 4096 = 64 MB: L2-resident
  printf("panels %d (%.0f MB)\n", panels_total, panels_total * 16384.0 / 1048576);
  const size_t bytes = (size_t)panels_total * kPanel;
  uint8_t* src;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMemset(src, 1, bytes));
  long long* cyc;
  CK(cudaMalloc(&cyc, 1024 * sizeof(long long)));
  const int smem = 1024 + kStages * kPanel + 24 * 1024;   // pad to the SpMM kernel's footprint: one CTA per SM
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d\n", sms);
  for (int csize : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / csize * csize);
    cfg.blockDim = dim3(64);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, bench, &cfg);
    printf("cluster size %2d: max co-resident clusters %d (%d CTAs) %s\n", csize, nclusters, nclusters * csize,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaGetLastError();
  }
  const int iters = 4000;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  {
    const int smem3 = 1024 + 12 * 16384;
    CK(cudaFuncSetAttribute(bench_two, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    for (int producers : {1, 2, 4})
      for (int per_iter : {1, 2, 4})
        for (int lin : {4096, 16384}) {
          float best = 1e30f;
          for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            bench_two<<<sms, 160, smem3>>>(lin, per_iter, producers, 12, iters, panels_total, cyc, (const uint8_t*)src);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
          }
          long long hc[1024];
          CK(cudaMemcpy(hc, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
          long long mx = 0;
          for (int i = 0; i < sms; ++i) mx = hc[i] > mx ? hc[i] : mx;
          printf("stages with their own barrier: %d producer warp(s), %d stage(s) issued back to back, %2d KB each, ring of 12: "
                 "%.0f cycles per stage, %.1f B/clk/SM\n", producers, per_iter, lin / 1024, (double)mx / iters, (double)lin * iters / mx);
        }
  }
  {
    CUtensorMap tm;
    const cuuint64_t dims[2] = {64, (cuuint64_t)panels_total * 128};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box_full[2] = {64, 128};
    encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box_full, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int smem2 = 1024 + 4 * (16384 + 32768);
    CK(cudaFuncSetAttribute(bench_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    for (int variant : {0})
    for (int S : {2, 4})
    for (int tensor_on = 0; tensor_on < 2; ++tensor_on)
      for (int lin : {0, 4096, 32768}) {
        if (!tensor_on && !lin) continue;
        if (variant == 2 && S != 4) continue;   // variant 2: the producer never waits for a free stage (pure issue rate)
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          CK(cudaEventRecord(e0));
          bench_mix<<<sms, 64, smem2>>>(tm, lin, tensor_on, iters, panels_total, cyc, (const uint8_t*)src, S, variant);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep > 0 && ms < best) best = ms;
        }
        long long hc[1024];
        CK(cudaMemcpy(hc, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < sms; ++i) mx = hc[i] > mx ? hc[i] : mx;
        const double per = (tensor_on ? 16384 : 0) + lin;
        printf("[%s] SpMM-like stage: tensor %d KB + linear %2d KB, %d stages: %.3f ms, %.2f TB/s, %.1f B/clk/SM, %.0f cycles per stage\n",
               variant == 0 ? "try_wait" : variant == 1 ? "test_wait" : "no empty wait", tensor_on ? 16 : 0, lin / 1024, S, best, per * sms * iters / best / 1e9, per * iters / (double)mx, (double)mx / iters);
      }
  }
  for (int csize : {1, 2, 4, 8}) {
    CUtensorMap tm_full, tm_slice;
    const cuuint64_t dims[2] = {64, (cuuint64_t)panels_total * 128};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box_full[2] = {64, 128};
    const cuuint32_t box_slice[2] = {64, (cuuint32_t)(128 / csize)};
    if (encode(&tm_full, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box_full, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        encode(&tm_slice, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box_slice, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      printf("tensor map encode failed\n");
      return 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(64);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(sms / csize * csize);
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, bench, &cfg) != cudaSuccess || nclusters == 0) { cudaGetLastError(); continue; }
    const int grid = std::min(nclusters * csize, sms / csize * csize);
    cfg.gridDim = dim3(grid);
    for (int mode = 0; mode < 8; ++mode) {
      if (csize == 1 && mode != 0 && mode != 3) continue;
      if (mode >= 5 && csize != 2) continue;   // tm_slice has 64-row boxes at cluster size 2
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, bench, tm_full, tm_slice, mode, csize, iters, panels_total, cyc, (const uint8_t*)src));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      long long hc[1024];
      CK(cudaMemcpy(hc, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = hc[i] > mx ? hc[i] : mx;
      const double landed = (double)grid * iters * kPanel;
      const double l2_reads = (mode == 0 || mode == 3 || mode >= 5) ? landed : landed / csize;
      printf("cluster %d grid %3d mode %d (%s): %.3f ms, landed %.2f TB/s (%.1f B/clk/SM), distinct L2 bytes %.2f TB/s, %lld cycles\n",
             csize, grid, mode, mode == 0 ? "unicast distinct" : mode == 1 ? "unicast shared " : mode == 2 ? "multicast      " : mode == 3 ? "linear bulk    " : mode == 4 ? "linear bulk mc " : mode == 5 ? "tensor + linear" : mode == 6 ? "2 x tensor 8 KB" : "2 x linear 8 KB", best,
             landed / best / 1e9, landed / grid / (double)mx, l2_reads / best / 1e9, mx);
    }
  }
  return 0;
}
