// Multi-GPU one-shot behind the C ABI (include/sparta_b200.h, sparta_vbr_spmm_multi).
//
// The reference has no multi-GPU path; its caller, test/cuda/cuda_multiply.cpp:
// 129-137, hands host pointers to one multiply routine.  This is that routine for the GPUs of one box:
// ONE process, one host thread per device while the shards are built, and
//   * A partitioned by contiguous block-row ranges balanced on modelled shard time
//     (sparta_partition_block_rows_modelled), every device uploads only its blocks;
//   * B uploaded ONCE (to device 0) and replicated with a single ncclBroadcast over NVLink;
//   * no communication during the multiply (block-rows are independent, vbr.cpp:342-368);
//   * C stays row-partitioned: every device copies its slab straight into the caller's C; with
//     gather_c the slabs are first all-gathered over NCCL (padded to the tallest) and the whole C
//     comes back from device 0, which is what a device-resident consumer would use.
// NCCL is loaded at run time (dlopen "libnccl.so.2") so that the library has no link dependency on it:
// a CPU-only box, or a process that already carries torch's NCCL, loads libsparta_b200 unchanged.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sparta_b200.h"

extern int sparta_internal_fail(int code, const std::string& msg);

namespace {

// the slice of the NCCL API used here (nccl.h: ncclResult_t is an int enum, 0 = success;
// ncclFloat32 = 7; communicators and the unique id are opaque)
typedef struct ncclComm* ncclComm_t;
typedef int (*CommInitAllFn)(ncclComm_t*, int, const int*);
typedef int (*CommDestroyFn)(ncclComm_t);
typedef int (*BroadcastFn)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
typedef int (*GroupFn)(void);
typedef const char* (*ErrStrFn)(int);
constexpr int kNcclFloat = 7;

struct Nccl {
  void* lib = nullptr;
  CommInitAllFn comm_init_all = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  BroadcastFn broadcast = nullptr;
  AllGatherFn all_gather = nullptr;
  GroupFn group_start = nullptr, group_end = nullptr;
  ErrStrFn err = nullptr;
  std::vector<ncclComm_t> comms;   // cached for the device count of the last call
  int n = 0;
};
std::mutex g_mutex;
Nccl g_nccl;

const char* load_nccl() {
  if (g_nccl.lib) return "";
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return "libnccl.so.2 not found (multi-GPU calls need NCCL)";
  g_nccl.comm_init_all = reinterpret_cast<CommInitAllFn>(dlsym(lib, "ncclCommInitAll"));
  g_nccl.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(lib, "ncclCommDestroy"));
  g_nccl.broadcast = reinterpret_cast<BroadcastFn>(dlsym(lib, "ncclBroadcast"));
  g_nccl.all_gather = reinterpret_cast<AllGatherFn>(dlsym(lib, "ncclAllGather"));
  g_nccl.group_start = reinterpret_cast<GroupFn>(dlsym(lib, "ncclGroupStart"));
  g_nccl.group_end = reinterpret_cast<GroupFn>(dlsym(lib, "ncclGroupEnd"));
  g_nccl.err = reinterpret_cast<ErrStrFn>(dlsym(lib, "ncclGetErrorString"));
  if (!g_nccl.comm_init_all || !g_nccl.comm_destroy || !g_nccl.broadcast || !g_nccl.all_gather || !g_nccl.group_start ||
      !g_nccl.group_end)
    return "libnccl.so.2 lacks a required symbol";
  g_nccl.lib = lib;
  return "";
}

}  // namespace

extern "C" int sparta_vbr_spmm_multi(int64_t rows, int64_t cols, int64_t block_rows, int64_t block_col_size,
                                     const int64_t* row_part, const int64_t* nzcount, const int64_t* jab,
                                     const float* mab, const float* B, int64_t ldb, int64_t n, float* C, int64_t ldc,
                                     int precision, int32_t n_gpus, int32_t gather_c, float* dt_ms, float* bcast_ms) {
  if (n_gpus <= 0 || !row_part || !B || !C || n <= 0 || ldb < cols || ldc < rows)
    return sparta_internal_fail(SPARTA_ERR_INVALID, "invalid multi-GPU request");
  if (sparta_device_count() < n_gpus)
    return sparta_internal_fail(SPARTA_ERR_NO_DEVICE, "fewer sm_100 devices visible than n_gpus");
  if (n_gpus == 1)
    return sparta_vbr_spmm(rows, cols, block_rows, block_col_size, row_part, nzcount, jab, mab, B, ldb, n, C, ldc, precision,
                           dt_ms);
  std::lock_guard<std::mutex> lock(g_mutex);
  const char* le = load_nccl();
  if (*le) return sparta_internal_fail(SPARTA_ERR_STATE, le);
  if (g_nccl.n != n_gpus) {
    for (ncclComm_t c : g_nccl.comms) g_nccl.comm_destroy(c);
    g_nccl.comms.assign(n_gpus, nullptr);
    std::vector<int> devs(n_gpus);
    for (int d = 0; d < n_gpus; ++d) devs[d] = d;
    const int r = g_nccl.comm_init_all(g_nccl.comms.data(), n_gpus, devs.data());
    if (r != 0) {
      g_nccl.comms.clear();
      g_nccl.n = 0;
      return sparta_internal_fail(SPARTA_ERR_CUDA, std::string("ncclCommInitAll: ") + (g_nccl.err ? g_nccl.err(r) : "error"));
    }
    g_nccl.n = n_gpus;
  }
  // shards balanced on modelled time
  sparta_options base;
  memset(&base, 0, sizeof(base));
  base.struct_size = sizeof(base);
  base.precision = precision;
  base.n_hint = static_cast<int32_t>(std::min<int64_t>(n, INT32_MAX));
  std::vector<int64_t> cuts(n_gpus + 1, 0);
  int rc = sparta_partition_block_rows_modelled(rows, cols, block_rows, block_col_size, row_part, nzcount, jab, n, &base,
                                                n_gpus, cuts.data());
  if (rc) return rc;
  // every device builds the handle of its shard on its own host thread (scheduling + upload overlap)
  std::vector<sparta_handle*> hs(n_gpus, nullptr);
  std::vector<int> rcs(n_gpus, 0);
  std::vector<std::string> errs(n_gpus);
  {
    std::vector<std::thread> th;
    for (int d = 0; d < n_gpus; ++d)
      th.emplace_back([&, d] {
        sparta_options o = base;
        o.device = d + 1;
        o.block_row_begin = cuts[d];
        o.block_row_end = cuts[d + 1];
        o.explicit_range = 1;
        rcs[d] = sparta_vbr_create(&hs[d], rows, cols, block_rows, block_col_size, row_part, nzcount, jab, mab, &o);
        if (rcs[d]) errs[d] = sparta_last_error();
      });
    for (auto& t : th) t.join();
  }
  auto cleanup = [&] { for (sparta_handle* h : hs) if (h) sparta_destroy(h); };
  for (int d = 0; d < n_gpus; ++d)
    if (rcs[d]) { cleanup(); return sparta_internal_fail(rcs[d], "shard " + std::to_string(d) + ": " + errs[d]); }
  // B: one upload, one broadcast
  std::vector<float*> dB(n_gpus, nullptr);
  const size_t b_elems = static_cast<size_t>(n) * cols;
  cudaError_t ce = cudaSuccess;
  for (int d = 0; d < n_gpus && ce == cudaSuccess; ++d) {
    ce = cudaSetDevice(d);
    if (ce == cudaSuccess) ce = cudaMallocAsync(reinterpret_cast<void**>(&dB[d]), b_elems * sizeof(float),
                                                static_cast<cudaStream_t>(sparta_stream(hs[d])));
  }
  cudaEvent_t b0 = nullptr, b1 = nullptr;
  if (ce == cudaSuccess) ce = cudaSetDevice(0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&b0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&b1);
  cudaStream_t s0 = static_cast<cudaStream_t>(sparta_stream(hs[0]));
  if (ce == cudaSuccess)
    ce = cudaMemcpy2DAsync(dB[0], cols * sizeof(float), B, ldb * sizeof(float), cols * sizeof(float), n,
                           cudaMemcpyHostToDevice, s0);
  if (ce == cudaSuccess) ce = cudaEventRecord(b0, s0);
  int nr = 0;
  if (ce == cudaSuccess) {
    g_nccl.group_start();
    for (int d = 0; d < n_gpus && nr == 0; ++d)
      nr = g_nccl.broadcast(dB[0], dB[d], b_elems, kNcclFloat, 0, g_nccl.comms[d], static_cast<cudaStream_t>(sparta_stream(hs[d])));
    const int ge = g_nccl.group_end();
    if (nr == 0) nr = ge;
  }
  if (ce == cudaSuccess && nr == 0) ce = cudaEventRecord(b1, s0);
  // multiply: every device on its own stream; no exchange
  std::vector<int64_t> slab_rows(n_gpus);
  int64_t tallest = 0;
  for (int d = 0; d < n_gpus; ++d) {
    slab_rows[d] = row_part[cuts[d + 1]] - row_part[cuts[d]];
    tallest = std::max(tallest, slab_rows[d]);
  }
  std::vector<float> times(n_gpus, 0.f);
  if (ce == cudaSuccess && nr == 0) {
    for (int d = 0; d < n_gpus && rc == 0; ++d)
      if (slab_rows[d] > 0) rc = sparta_set_B(hs[d], dB[d], cols, n, 1);
    std::vector<std::thread> th;
    std::vector<int> rr(n_gpus, 0);
    for (int d = 0; d < n_gpus && rc == 0; ++d)
      th.emplace_back([&, d] { if (slab_rows[d] > 0) rr[d] = sparta_run(hs[d], &times[d]); });
    for (auto& t : th) t.join();
    for (int d = 0; d < n_gpus; ++d) if (rr[d] && !rc) rc = rr[d];
  }
  // C back
  if (ce == cudaSuccess && nr == 0 && rc == 0) {
    if (!gather_c) {
      // the slab of rows [r0, r0 + h) of a column-major C is a strided region: one 2-D copy per device
      // (in parallel: every GPU has its own PCIe link)
      std::vector<std::thread> th;
      std::vector<int> rr(n_gpus, 0);
      for (int d = 0; d < n_gpus; ++d)
        th.emplace_back([&, d] {
          if (slab_rows[d] > 0) rr[d] = sparta_get_C(hs[d], C + (row_part[cuts[d]] - row_part[0]), ldc, 0);
          if (rr[d]) errs[d] = sparta_last_error();
        });
      for (auto& t : th) t.join();
      for (int d = 0; d < n_gpus; ++d)
        if (rr[d] && !rc) { rc = rr[d]; sparta_internal_fail(rc, errs[d]); }
    } else {
      // all-gather of the slabs padded to the tallest, then the whole C from device 0
      std::vector<float*> slab(n_gpus, nullptr), all(n_gpus, nullptr);
      const size_t per = static_cast<size_t>(n) * tallest;
      for (int d = 0; d < n_gpus && ce == cudaSuccess; ++d) {
        cudaStream_t s = static_cast<cudaStream_t>(sparta_stream(hs[d]));
        ce = cudaSetDevice(d);
        if (ce == cudaSuccess) ce = cudaMallocAsync(reinterpret_cast<void**>(&slab[d]), per * sizeof(float), s);
        if (ce == cudaSuccess) ce = cudaMallocAsync(reinterpret_cast<void**>(&all[d]), per * n_gpus * sizeof(float), s);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(slab[d], 0, per * sizeof(float), s);
        if (ce == cudaSuccess && slab_rows[d] > 0) {
          ce = cudaStreamSynchronize(s);
          if (ce == cudaSuccess && sparta_get_C(hs[d], slab[d], tallest, 1)) ce = cudaErrorUnknown;   // [n][tallest]
        }
      }
      if (ce == cudaSuccess) {
        g_nccl.group_start();
        for (int d = 0; d < n_gpus && nr == 0; ++d)
          nr = g_nccl.all_gather(slab[d], all[d], per, kNcclFloat, g_nccl.comms[d], static_cast<cudaStream_t>(sparta_stream(hs[d])));
        const int ge = g_nccl.group_end();
        if (nr == 0) nr = ge;
      }
      if (ce == cudaSuccess && nr == 0) ce = cudaSetDevice(0);
      for (int d = 0; d < n_gpus && ce == cudaSuccess && nr == 0; ++d)
        if (slab_rows[d] > 0)
          ce = cudaMemcpy2DAsync(C + (row_part[cuts[d]] - row_part[0]), ldc * sizeof(float), all[0] + per * d,
                                 tallest * sizeof(float), slab_rows[d] * sizeof(float), n, cudaMemcpyDeviceToHost, s0);
      for (int d = 0; d < n_gpus; ++d) {
        cudaSetDevice(d);
        cudaStream_t s = static_cast<cudaStream_t>(sparta_stream(hs[d]));
        cudaStreamSynchronize(s);
        if (slab[d]) cudaFreeAsync(slab[d], s);
        if (all[d]) cudaFreeAsync(all[d], s);
      }
    }
  }
  float bms = 0.f;
  if (ce == cudaSuccess && nr == 0) {
    cudaSetDevice(0);
    cudaStreamSynchronize(s0);
    if (cudaEventElapsedTime(&bms, b0, b1) != cudaSuccess) { bms = 0.f; cudaGetLastError(); }
  }
  for (int d = 0; d < n_gpus; ++d) {
    cudaSetDevice(d);
    if (dB[d]) cudaFreeAsync(dB[d], static_cast<cudaStream_t>(sparta_stream(hs[d])));
  }
  if (b0) cudaEventDestroy(b0);
  if (b1) cudaEventDestroy(b1);
  const std::string keep = rc ? sparta_last_error() : "";
  cleanup();
  cudaSetDevice(0);
  if (ce != cudaSuccess) return sparta_internal_fail(SPARTA_ERR_CUDA, std::string("multi-GPU one-shot: ") + cudaGetErrorString(ce));
  if (nr != 0) return sparta_internal_fail(SPARTA_ERR_CUDA, std::string("NCCL: ") + (g_nccl.err ? g_nccl.err(nr) : "error"));
  if (rc) return sparta_internal_fail(rc, keep);
  if (dt_ms) *dt_ms = *std::max_element(times.begin(), times.end());
  if (bcast_ms) *bcast_ms = bms;
  return SPARTA_OK;
}
