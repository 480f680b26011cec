// extern "C" boundary of libsparta_b200 (include/sparta_b200.h).
//
// Host-side responsibilities mirror what each reference multiply routine does
// around its GEMM loop (src/cuda/cuda_utilities.cpp:91-105, 199-208): allocate,
// upload A and B, time the compute with CUDA events, download C.  Here the
// device state lives in a handle so warm-up and repetitions reuse it.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/sparta_b200.h"
#include "blocking.h"
#include "csr_kernel.h"
#include "host_formats.h"
#include "pack_kernels.h"
#include "schedule.h"
#include "spmm_kernel.h"

using namespace sparta;

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
static int fail_cuda(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return SPARTA_ERR_CUDA;
}
// No C++ exception may cross the extern "C" boundary (ctypes, the reference's C++11 caller): the entry
// points that allocate on the host run their bodies through this.
template <class F>
static int guarded(F body) {
  try {
    return body();
  } catch (const std::bad_alloc&) {
    return fail(SPARTA_ERR_INVALID, "out of host memory");
  } catch (const std::exception& e) {
    return fail(SPARTA_ERR_INVALID, std::string("internal error: ") + e.what());
  } catch (...) {
    return fail(SPARTA_ERR_INVALID, "internal error");
  }
}

// for the other translation units of the library (multi_gpu.cu)
int sparta_internal_fail(int code, const std::string& msg) { return fail(code, msg); }
#define CU_TRY(call)                                        \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) return fail_cuda(e_, #call);     \
  } while (0)

// ---- device memory: the stream-ordered allocator with a retained pool ------------------------
// Every device buffer comes from the device's default memory pool (cudaMallocAsync) whose release
// threshold is raised once per device so that freed blocks stay cached: the one-shot calls
// (upload, multiply, download per call, like the reference's routines) then pay cudaMalloc /
// cudaFree -- tens of milliseconds for multi-GB buffers, and a device-wide synchronisation
// each -- only on their first use.  sparta_release_workspace() hands the cache back.
static std::mutex g_pool_mutex;
static bool g_pool_ready[64] = {};

static cudaError_t pool_prepare(int dev) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (dev < 0 || dev >= 64 || g_pool_ready[dev]) return cudaSuccess;
  cudaMemPool_t pool;
  cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, dev);
  if (e != cudaSuccess) return e;
  uint64_t keep = UINT64_MAX;
  e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  if (e == cudaSuccess) g_pool_ready[dev] = true;
  return e;
}
static cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t s) {
  *p = nullptr;
  return cudaMallocAsync(p, std::max<size_t>(bytes, 16), s);
}
template <class T>
static cudaError_t dev_alloc(T** p, size_t bytes, cudaStream_t s) {
  return dev_alloc(reinterpret_cast<void**>(p), bytes, s);
}
template <class T>
static void dev_free(T*& p, cudaStream_t s) {
  if (p) cudaFreeAsync(const_cast<void*>(reinterpret_cast<const void*>(p)), s);
  p = nullptr;
}

// ---- pinned host staging, cached like the device pool ------------------------------------------
// What the library itself builds on the host and sends up (the nonzeros of a CSR-built handle, the
// schedule arrays, the pack jobs: ~60 MB for BASELINE config #3) used to leave from pageable memory:
// the driver stages such copies through its own bounce buffer at a few GB/s and the call blocks until
// the stream reaches it.  Blocks of page-locked memory are kept in a process-wide free list (first use
// pays cudaHostAlloc, like the device pool pays cudaMalloc); sparta_release_workspace() frees them.
struct PinnedCache {
  std::mutex m;
  std::vector<std::pair<void*, size_t>> free_list;
  std::unordered_map<void*, size_t> live;
};
static PinnedCache& pinned_cache() {
  static PinnedCache* c = new PinnedCache();   // never destroyed: handles may outlive static destructors
  return *c;
}
static void* pinned_acquire(size_t bytes) {
  if (getenv("SPARTA_PAGEABLE_STAGING")) return nullptr;
  // size classes: powers of two from 64 KB to 64 MB, multiples of 64 MB above.  A request is served only by a
  // block of ITS class, so a call that repeats (the one-shot routines) finds every block it used the last time
  // whatever order its requests come in -- cudaHostAlloc synchronises the device and costs milliseconds
  // (hundreds with several processes on the box), it must not recur.
  size_t want = size_t{1} << 16;
  while (want < bytes && want < (size_t{1} << 26)) want <<= 1;
  if (want < bytes) want = (bytes + (size_t{1} << 26) - 1) >> 26 << 26;
  PinnedCache& c = pinned_cache();
  {
    std::lock_guard<std::mutex> lock(c.m);
    for (size_t i = 0; i < c.free_list.size(); ++i)
      if (c.free_list[i].second == want) {
        const auto blk = c.free_list[i];
        c.free_list.erase(c.free_list.begin() + static_cast<std::ptrdiff_t>(i));
        c.live[blk.first] = blk.second;
        return blk.first;
      }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(c.m);
  c.live[p] = want;
  return p;
}
static void pinned_release(void* p) {
  if (!p) return;
  PinnedCache& c = pinned_cache();
  std::lock_guard<std::mutex> lock(c.m);
  auto it = c.live.find(p);
  if (it == c.live.end()) return;
  c.free_list.emplace_back(p, it->second);
  c.live.erase(it);
}
static void pinned_trim() {
  PinnedCache& c = pinned_cache();
  std::vector<std::pair<void*, size_t>> blocks;
  {
    std::lock_guard<std::mutex> lock(c.m);
    blocks.swap(c.free_list);
  }
  for (auto& b : blocks) cudaFreeHost(b.first);
  cudaGetLastError();
}
// RawBuf hooks (host_formats.h): page-locked when the cache can supply it, else the C heap
static void* pinned_or_malloc(size_t bytes, bool* pinned) {
  void* p = pinned_acquire(bytes);
  *pinned = p != nullptr;
  return p ? p : malloc(bytes ? bytes : 1);
}
static void pinned_or_free(void* p, bool pinned) {
  if (pinned) pinned_release(p); else free(p);
}

struct sparta_host_vbr { HostVBR v; };
struct sparta_host_bell { HostBell b; };

struct sparta_plan {
  Structure st;
  Assignment as;
  ScheduleOptions sopt;
  int64_t cols = 0, block_rows = 0, n = 0;
  int panel_stages = 4, a_ring_bytes = 0;
};

struct sparta_handle {
  bool chain_ok = false;       // the last kernel enqueued on the stream is this handle's own SpMM launch and nothing
                               // it reads before its epilogue has changed since (cleared by set_B): the next launch
                               // may overlap its tail (spmm_launch, overlap_previous)
  std::vector<void*> staged;   // pinned blocks the stream may still be reading (released after a synchronisation)
  int device = 0;
  int kind = 0;   // 0: block-sparse (VBR / Blocked-ELL) on the tcgen05 kernel, 1: CSR gather kernel
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t up0 = nullptr, up1 = nullptr;   // around the last upload (create / set_B)
  ScheduleOptions sopt;
  int b_row_major = 0, c_row_major = 0, accumulate = 0;
  int panel_stages = 4, a_ring_bytes = 0;
  int a_slot_bytes = 0;   // > 0: fixed-slot pipeline (spmm_kernel.h)
  int producers = 1;
  int64_t cols = 0, block_rows = 0, w = 0;
  int64_t rows = 0;   // C rows of the shard
  Structure st;   // jobs are dropped after packing
  Assignment as;
  Segment* d_segs = nullptr;
  SuperRow* d_srows = nullptr;
  Chunk* d_chunks = nullptr;
  uint32_t* d_tables = nullptr;
  uint8_t* d_a = nullptr;
  Item* d_items = nullptr;
  int32_t* d_cta_ptr = nullptr;
  int32_t* d_cta_items = nullptr;
  ZeroJob* d_zero_jobs = nullptr;
  unsigned long long* d_sync = nullptr;   // grid-wide counter of the in-kernel zeroing (SpmmParams)
  unsigned long long sync_total = 0;      // CTAs that have bumped it over all launches so far
  // CSR handles
  int64_t* d_rowptr = nullptr;
  int32_t* d_colind = nullptr;
  float* d_val = nullptr;
  int32_t* d_row_order = nullptr;
  int64_t nnz = 0, heavy_rows = 0;
  // hybrid VBR handles: the block-rows of at most gather_max_height rows run on the gather kernel;
  // their nonzeros live in the CSR arrays above (rowptr over ALL rows of the shard, the work list
  // row_order holds the gather rows only), B is kept a second time with its ROWS contiguous
  int64_t gather_rows = 0, gather_work = 0;   // rows routed to the gather kernel / of them with nonzeros
  struct GatherPassDev { int64_t* beg = nullptr; int64_t* end = nullptr; int32_t* order = nullptr; int64_t work = 0, heavy = 0; };
  std::vector<GatherPassDev> gather_passes;   // one launch per range of k (GatherPart::Pass)
  void* d_B2 = nullptr;
  size_t b2_cap = 0;
  int64_t ldn2 = 0;
  void* d_B = nullptr;
  size_t b_cap = 0;
  int64_t ldk = 0, n = 0;
  float* d_C = nullptr;
  size_t c_cap = 0;
  int64_t ldc = 0;
  int64_t launches = 0;
};

static void resolve_options(const sparta_options* in, sparta_options* o) {
  memset(o, 0, sizeof(*o));
  if (in) {
    size_t sz = in->struct_size > 0 ? static_cast<size_t>(in->struct_size) : sizeof(*o);
    memcpy(o, in, std::min(sz, sizeof(*o)));
  }
  if (o->seg_rows == 0) o->seg_rows = 64;
  if (o->acc_cols == 0) o->acc_cols = 512;
  if (o->l2_slab_mb <= 0) o->l2_slab_mb = 160;
  if (o->panel_stages == 0) o->panel_stages = 5;
}

// One past the last unit of the shard.  A zeroed options struct selects everything; with
// explicit_range the pair [begin, end) is literal, so begin == end is an EMPTY shard (a balanced
// partition may hand a rank nothing) instead of silently meaning "all".
static int64_t range_end(const sparta_options& o, int64_t total) {
  if (o.explicit_range) return o.block_row_end;
  return o.block_row_end > 0 ? o.block_row_end : total;
}

// Persistent grid size of the host-only planning entry points (plans, modelled partitions): the
// SM count of the current device when there is one, else the B200's 148.
static int default_grid_ctas() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
    cached = sms;
  else {
    cudaGetLastError();
    cached = 148;
  }
  return cached;
}

static int ring_bytes_for(int panel_stages) {
  int b = kSmemMax - 1024 - kSmemCtrlBytes - panel_stages * kPanelBytes;
  return b / 1024 * 1024;
}

// Shared-memory pipeline of a handle (spmm_kernel.h): fixed slots -- every stage owns room for the
// largest chunk's A images -- when at least 3 such stages fit (they do unless a single-CTA handle
// has 512-row chunks), else the byte ring.  Slots make the producer's per-chunk work short enough
// for two copy warps to matter; chunks of 1-2 members (ER-like matrices, variable-height blockings)
// are bound by how fast stages can be STARTED, not by bytes.
static void pick_pipeline(uint32_t max_chunk_bytes, const sparta_options& o, int tiles, int* stages, int* ring_bytes,
                          int* slot_bytes, int* producers) {
  const int avail = kSmemMax - 1024 - kSmemCtrlBytes;
  const int slot = static_cast<int>((std::max<uint32_t>(max_chunk_bytes, 1024) + 1023) / 1024 * 1024);
  const int fit = std::min(kMaxPanelStages, avail / (tiles * kPanelBytes + slot));
  const bool slots = tiles > 1 || (o.pipeline == 2 ? fit >= 2 : (o.pipeline == 1 ? false : fit >= 3));
  if (slots) {
    *stages = fit;
    *slot_bytes = slot;
    *ring_bytes = fit * slot;
    *producers = o.copy_warps == 1 ? 1 : 2;
  } else {
    *slot_bytes = 0;
    *producers = 1;   // stages and ring as set from the options
  }
}

// ---- BlockRows builders (format adapters) ---------------------------------

static const char* blockrows_from_vbr(int64_t block_rows, int64_t w, const int64_t* row_part,
                                      const int64_t* nzcount, const int64_t* jab, int64_t lo,
                                      int64_t hi, BlockRows* br, int64_t* src_lo,
                                      int64_t* src_hi) {
  if (lo < 0 || hi > block_rows || lo > hi) return "block-row range out of bounds";
  br->w = w;
  int64_t jab_off = 0, mab_off = 0;
  for (int64_t b = 0; b < lo; ++b) {
    const int64_t H = row_part[b + 1] - row_part[b];
    if (H < 0 || nzcount[b] < 0) return "row_part must be non-decreasing and nzcount non-negative";
    jab_off += nzcount[b];
    mab_off += nzcount[b] * H * w;
  }
  *src_lo = mab_off;
  br->ptr.push_back(0);
  for (int64_t b = lo; b < hi; ++b) {
    const int64_t H = row_part[b + 1] - row_part[b];
    if (H < 0 || nzcount[b] < 0) return "row_part must be non-decreasing and nzcount non-negative";
    br->row0.push_back(row_part[b] - row_part[lo]);
    br->height.push_back(H);
    br->rs.push_back(1);   // blocks are column-major with ld = H (vbr.cpp:224)
    br->ks.push_back(H);
    for (int64_t q = 0; q < nzcount[b]; ++q) {
      const int64_t jb = jab[jab_off + q];
      if (q > 0 && jb <= jab[jab_off + q - 1]) return "jab must be strictly ascending inside a block-row";
      if (jb < 0) return "negative column-block index";
      br->col.push_back(jb);
      br->src.push_back(mab_off - *src_lo + q * H * w);
    }
    br->ptr.push_back(static_cast<int64_t>(br->col.size()));
    jab_off += nzcount[b];
    mab_off += nzcount[b] * H * w;
  }
  *src_hi = mab_off;
  return "";
}

// A TRANSPOSED, for the inverted product C = B*A (cublas_blockmat_multiplyBA,
// cuda_utilities.cpp:553-721): the block-rows of the operand are A's column blocks [lo, hi) --
// `w` output rows each, the last one cols - jb*w -- and its blocks along k are A's block-rows,
// k = A's rows in blocked order (row_part), extent = the block-row's height.  Source element (r, k)
// of the transposed block (ib, jb) is mab[block + k + r*h]: row stride h, k stride 1.
static const char* blockrows_from_vbr_transposed(int64_t cols, int64_t block_rows, int64_t w,
                                                 const int64_t* row_part, const int64_t* nzcount,
                                                 const int64_t* jab, int64_t lo, int64_t hi, BlockRows* br,
                                                 int64_t* src_hi) {
  const int64_t bc = (cols - 1) / w + 1;
  if (lo < 0 || hi > bc || lo > hi) return "column-block range out of bounds";
  br->w = w;
  const int64_t nb = hi - lo;
  std::vector<int64_t> count(nb + 1, 0);
  int64_t q = 0, mab_off = 0;
  for (int64_t b = 0; b < block_rows; ++b) {
    const int64_t H = row_part[b + 1] - row_part[b];
    if (H < 0 || nzcount[b] < 0) return "row_part must be non-decreasing and nzcount non-negative";
    for (int64_t t = 0; t < nzcount[b]; ++t, ++q) {
      const int64_t jb = jab[q];
      if (t > 0 && jb <= jab[q - 1]) return "jab must be strictly ascending inside a block-row";
      if (jb < 0 || jb >= bc) return "jab entry beyond the last column block";
      if (jb >= lo && jb < hi && H > 0) ++count[jb - lo + 1];
    }
    mab_off += nzcount[b] * H * w;
  }
  *src_hi = mab_off;
  for (int64_t j = 0; j < nb; ++j) count[j + 1] += count[j];
  br->ptr = count;
  const int64_t total = count[nb];
  br->col.resize(total); br->src.resize(total);
  br->blk_k0.resize(total); br->blk_kw.resize(total); br->blk_rs.resize(total); br->blk_ks.assign(total, 1);
  std::vector<int64_t> fill(count.begin(), count.end() - 1);
  q = 0; mab_off = 0;
  for (int64_t b = 0; b < block_rows; ++b) {      // ascending b => ascending k inside every list
    const int64_t H = row_part[b + 1] - row_part[b];
    for (int64_t t = 0; t < nzcount[b]; ++t, ++q) {
      const int64_t jb = jab[q];
      if (jb >= lo && jb < hi && H > 0) {
        const int64_t at = fill[jb - lo]++;
        br->col[at] = b;
        br->src[at] = mab_off + t * H * w;
        br->blk_k0[at] = row_part[b];
        br->blk_kw[at] = H;
        br->blk_rs[at] = H;
      }
    }
    mab_off += nzcount[b] * H * w;
  }
  for (int64_t j = lo; j < hi; ++j) {
    br->row0.push_back((j - lo) * w);
    br->height.push_back(std::min(w, cols - j * w));
    br->rs.push_back(0);
    br->ks.push_back(0);
  }
  return "";
}

static const char* blockrows_from_bell(int64_t bs, int64_t ind_rows, int64_t ind_cols,
                                       const int64_t* ind, int64_t lo, int64_t hi, BlockRows* br,
                                       int64_t* src_lo, int64_t* src_hi) {
  if (lo < 0 || hi > ind_rows || lo > hi) return "block-row range out of bounds";
  br->w = bs;
  const int64_t val_cols = ind_cols * bs;  // ellValue_cols (cuda_utilities.cpp:1681)
  *src_lo = lo * bs * val_cols;
  *src_hi = hi * bs * val_cols;
  br->ptr.push_back(0);
  std::vector<std::pair<int64_t, int64_t>> row;
  for (int64_t i = lo; i < hi; ++i) {
    br->row0.push_back((i - lo) * bs);
    br->height.push_back(bs);
    br->rs.push_back(val_cols);  // values are row-major rows x ellValue_cols (:1699-1707)
    br->ks.push_back(1);
    row.clear();
    for (int64_t s = 0; s < ind_cols; ++s) {
      const int64_t jb = ind[i * ind_cols + s];
      if (jb < 0) continue;  // -1 = padding block (:1693)
      row.emplace_back(jb, (i - lo) * bs * val_cols + s * bs);
    }
    std::sort(row.begin(), row.end());
    for (size_t t = 0; t < row.size(); ++t) {
      if (t > 0 && row[t].first == row[t - 1].first) return "duplicate column block in an ELL row";
      br->col.push_back(row[t].first);
      br->src.push_back(row[t].second);
    }
    br->ptr.push_back(static_cast<int64_t>(br->col.size()));
  }
  return "";
}

// ---- short block-rows: off the tensor cores ----------------------------------------------------
// A tcgen05.mma needs 16 rows of N (8 in single-CTA mode): a block-row of height 1 is 94 % padding
// and pulls a 16 KB panel of B through the L2 to use one row of it per nonzero.  Variable-height
// blockings (-a 3 / -a 4) leave most block-rows that short (BASELINE config #4: 127 586 of 137 137
// block-rows have height 1 and hold 63 % of the stored elements).  Those block-rows are taken out
// of the tile schedule and handed, as the nonzeros of their blocks, to the gather kernel of the
// family (csr_kernel.cu); everything else is unchanged.  Stored zeros contribute nothing to
// VBR::multiply (vbr.cpp:358-363), so the product is the same.
struct GatherPart {
  std::vector<int64_t> rowptr;    // [shard rows + 1]
  std::vector<int32_t> colind;
  std::vector<float> val;
  // The nonzeros are walked in PASSES over ranges of k (columns of A = rows of B): one launch per pass,
  // so that the part of B a pass reads -- (cols / passes) rows x 256 columns per column tile -- stays
  // L2-resident; with one pass at 2^18 columns a tile's slab of B is 268 MB and every gathered row comes
  // from HBM.  Pass 0 lists every gather row that has nonzeros anywhere (it WRITES the row, zeros
  // included); later passes add to C and list only the rows with nonzeros in their range.
  struct Pass {
    std::vector<int64_t> beg, end;   // [shard rows]: the row's entries in this pass's range of k
    std::vector<int32_t> order;      // work list, longest first
    int64_t heavy = 0;               // leading entries of `order` longer than kCsrHeavyNnz
  };
  std::vector<Pass> passes;
  int64_t rows = 0;               // rows of the block-rows taken out
  int64_t nztot = 0, blocks = 0;  // their stored elements / nonzero blocks (FLOP accounting, vbr.cpp:232)
};

// The ranges of k of a gather part (GatherPart::Pass): `live` = the gather rows that hold nonzeros.
static void build_gather_passes(GatherPart* g, const std::vector<int32_t>& live, int64_t total_rows, int64_t cols,
                                int esize_b, int force_passes) {
  // a pass's slab of B (range x 256 columns) within ~48 MB
  const double slab = static_cast<double>(cols) * 256.0 * esize_b;
  // Measured at BASELINE config #4 (2^18 columns, fp32 rows of B, 70 nonzeros per gather row): 1 pass
  // 38.0 ms, 6 passes 102.9, 8 passes 126.1 -- the rows' segments get too short to amortise a warp's
  // fixed work and every pass after the first read-modify-writes its rows of C, so one pass is the default
  // although a column tile's slab of B (268 MB there) no longer fits the L2.
  (void)slab;
  const int P = force_passes > 0 ? std::min(force_passes, 64) : 1;
  const int64_t range = ((cols + P - 1) / P + 63) / 64 * 64;
  g->passes.assign(P, GatherPart::Pass());
  for (int pi = 0; pi < P; ++pi) {
    GatherPart::Pass& ps = g->passes[pi];
    ps.beg.assign(static_cast<size_t>(total_rows), 0);
    ps.end.assign(static_cast<size_t>(total_rows), 0);
  }
  for (int32_t r : live) {
    int64_t at = g->rowptr[r];
    const int64_t stop = g->rowptr[r + 1];
    for (int pi = 0; pi < P; ++pi) {
      const int64_t k_hi = (pi + 1 == P) ? cols : (pi + 1) * range;
      const int64_t first = at;
      while (at < stop && g->colind[at] < k_hi) ++at;     // columns ascend inside a row
      g->passes[pi].beg[r] = first;
      g->passes[pi].end[r] = at;
    }
  }
  for (int pi = 0; pi < P; ++pi) {
    GatherPart::Pass& ps = g->passes[pi];
    for (int32_t r : live)
      if (pi == 0 || ps.end[r] > ps.beg[r]) ps.order.push_back(r);
    std::stable_sort(ps.order.begin(), ps.order.end(),
                     [&](int32_t a, int32_t c) { return ps.end[a] - ps.beg[a] > ps.end[c] - ps.beg[c]; });
    ps.heavy = 0;
    while (ps.heavy < static_cast<int64_t>(ps.order.size()) &&
           ps.end[ps.order[ps.heavy]] - ps.beg[ps.order[ps.heavy]] > kCsrHeavyNnz)
      ++ps.heavy;
  }
}

static bool split_short_block_rows(const BlockRows& br, int max_height, const float* src, int64_t cols,
                                   int esize_b, int force_passes, BlockRows* tall, GatherPart* g) {
  const int64_t nb = br.count();
  if (max_height <= 0 || nb == 0 || !br.blk_k0.empty() || !br.sub_ptr.empty()) return false;
  std::vector<int64_t> shorts;
  for (int64_t b = 0; b < nb; ++b)
    if (br.height[b] > 0 && br.height[b] <= max_height && br.ptr[b + 1] > br.ptr[b]) shorts.push_back(b);
  if (shorts.empty()) return false;
  int64_t total_rows = 0;
  for (int64_t b = 0; b < nb; ++b) total_rows = std::max(total_rows, br.row0[b] + br.height[b]);
  if (total_rows > INT32_MAX || cols > INT32_MAX) return false;
  // the tall view: same rows of C, same sources, the short block-rows dropped
  BlockRows t;
  t.w = br.w;
  t.ptr.push_back(0);
  {
    size_t si = 0;
    for (int64_t b = 0; b < nb; ++b) {
      if (si < shorts.size() && shorts[si] == b) { ++si; continue; }
      t.row0.push_back(br.row0[b]); t.height.push_back(br.height[b]); t.rs.push_back(br.rs[b]); t.ks.push_back(br.ks[b]);
      for (int64_t q = br.ptr[b]; q < br.ptr[b + 1]; ++q) { t.col.push_back(br.col[q]); t.src.push_back(br.src[q]); }
      t.ptr.push_back(static_cast<int64_t>(t.col.size()));
    }
  }
  // nonzeros of the short block-rows, scanned by all host threads over contiguous ranges of them
  const int T = static_cast<int>(std::max<size_t>(1, std::min<size_t>(host_thread_budget(32), shorts.size() / 64 + 1)));
  std::vector<std::vector<int32_t>> t_col(T);
  std::vector<std::vector<float>> t_val(T);
  std::vector<int64_t> row_nnz(static_cast<size_t>(total_rows), 0);
  auto work = [&](int tid) {
    const size_t lo = shorts.size() * tid / T, hi = shorts.size() * (tid + 1) / T;
    std::vector<int32_t>& cv = t_col[tid];
    std::vector<float>& vv = t_val[tid];
    for (size_t si = lo; si < hi; ++si) {
      const int64_t b = shorts[si];
      const int64_t H = br.height[b], rs = br.rs[b], ks = br.ks[b];
      for (int64_t r = 0; r < H; ++r) {
        int64_t cnt = 0;
        for (int64_t q = br.ptr[b]; q < br.ptr[b + 1]; ++q) {
          const float* blk = src + br.src[q] + r * rs;
          const int64_t k0 = br.col[q] * br.w;
          const int64_t kw = std::min<int64_t>(br.w, cols - k0);
          for (int64_t c = 0; c < kw; ++c) {
            const float x = blk[c * ks];
            if (x != 0.0f) { cv.push_back(static_cast<int32_t>(k0 + c)); vv.push_back(x); ++cnt; }
          }
        }
        row_nnz[static_cast<size_t>(br.row0[b] + r)] = cnt;
      }
    }
  };
  std::vector<std::thread> th;
  for (int tid = 1; tid < T; ++tid) th.emplace_back(work, tid);
  work(0);
  for (auto& x : th) x.join();
  g->rowptr.assign(static_cast<size_t>(total_rows) + 1, 0);
  for (int64_t r = 0; r < total_rows; ++r) g->rowptr[r + 1] = g->rowptr[r] + row_nnz[r];
  const int64_t nnz = g->rowptr[total_rows];
  g->colind.resize(static_cast<size_t>(nnz));
  g->val.resize(static_cast<size_t>(nnz));
  {
    // thread t's entries are the rows of its block-rows in order: copy them behind each other
    // (the short block-rows are visited in ascending row order, and rowptr was built the same way)
    int64_t at = 0;
    for (int tid = 0; tid < T; ++tid) {
      std::copy(t_col[tid].begin(), t_col[tid].end(), g->colind.begin() + at);
      std::copy(t_val[tid].begin(), t_val[tid].end(), g->val.begin() + at);
      at += static_cast<int64_t>(t_col[tid].size());
    }
  }
  g->rows = 0; g->nztot = 0; g->blocks = 0;
  std::vector<int32_t> live;      // gather rows that hold nonzeros
  for (int64_t b : shorts) {
    g->rows += br.height[b];
    g->blocks += br.ptr[b + 1] - br.ptr[b];
    g->nztot += (br.ptr[b + 1] - br.ptr[b]) * br.height[b] * br.w;
    for (int64_t r = 0; r < br.height[b]; ++r)
      if (row_nnz[static_cast<size_t>(br.row0[b] + r)] > 0) live.push_back(static_cast<int32_t>(br.row0[b] + r));
  }
  build_gather_passes(g, live, total_rows, cols, esize_b, force_passes);
  *tall = std::move(t);
  return true;
}


// The same split on the index arrays alone (no values at hand: partition requests): the tall view
// and an ESTIMATE of the gather rows' nonzeros -- blocks of a short block-row of a clustered sparse
// matrix hold little more than one nonzero per row.
static bool split_short_view(const BlockRows& br, int max_height, BlockRows* tall, double* est_nnz) {
  const int64_t nb = br.count();
  *est_nnz = 0;
  if (max_height <= 0 || nb == 0 || !br.blk_k0.empty() || !br.sub_ptr.empty()) return false;
  BlockRows t;
  t.w = br.w;
  t.ptr.push_back(0);
  bool any = false;
  for (int64_t b = 0; b < nb; ++b) {
    if (br.height[b] > 0 && br.height[b] <= max_height && br.ptr[b + 1] > br.ptr[b]) {
      *est_nnz += 1.2 * static_cast<double>(br.ptr[b + 1] - br.ptr[b]) * static_cast<double>(br.height[b]);
      any = true;
      continue;
    }
    t.row0.push_back(br.row0[b]); t.height.push_back(br.height[b]); t.rs.push_back(br.rs[b]); t.ks.push_back(br.ks[b]);
    for (int64_t q = br.ptr[b]; q < br.ptr[b + 1]; ++q) { t.col.push_back(br.col[q]); t.src.push_back(br.src[q]); }
    t.ptr.push_back(static_cast<int64_t>(t.col.size()));
  }
  if (any) *tall = std::move(t);
  return any;
}

// The tile schedule, with the number of column tiles per work item (sparta_options::wide_tiles) chosen
// here when the caller left it open.  Measured on B200 (profiles/r2_wide_items.md): two tiles per item
// are 5-14 % faster than one on every bf16 shape tried once n spans 8 tile widths (2048 columns for CTA
// pairs: config #3 3.09 -> 2.85 ms), neutral (-5..+6 %) at 4 widths, and at 2-4 widths they pay only
// when the member block-rows of a super-row rarely share a column block (ER-like lists: a B panel feeds
// one block, and the rate at which stages can be STARTED, not their bytes, bounds the pipeline --
// config #2 0.122 -> 0.099 ms).  Four tiles per item were slower everywhere but there (and not the
// best there either).  So: n_hint >= 6 tile widths -> 2 tiles; 2..6 widths -> 2 tiles when a first build
// with one tile shows fewer than 1.6 members per chunk; unknown or small n -> 1.  Schedules that need
// bounded accumulation chains (tf32 by default) keep one tile: the master accumulators need the room.
static const char* build_structure_choosing_tiles(const BlockRows& view, const sparta_options& o, ScheduleOptions* so,
                                                  Structure* st) {
  const int tile_w = so->pair ? 2 * kTileJ : kTileJ;
  auto build_with = [&](int tiles, Structure* out) {
    ScheduleOptions s2 = *so;
    s2.tiles = tiles;
    if (tiles > 1) s2.acc_cols = 512 / tiles;
    const char* e = build_structure(view, s2, out);      // (falls back to one tile if chains must be bounded)
    if (!*e && out->tiles == tiles && tiles > 1) so->tiles = tiles, so->acc_cols = s2.acc_cols;
    return e;
  };
  if (o.wide_tiles == 2 || o.wide_tiles == 4) return build_with(o.wide_tiles, st);
  if (so->acc_cols != 256 && so->acc_cols != 512) return "acc_cols must be 256 or 512";
  const bool open = o.wide_tiles == 0 && so->acc_cols == 512 && o.n_hint >= 2 * tile_w;
  if (open && o.n_hint >= 6 * tile_w) return build_with(2, st);
  const char* e = build_structure(view, *so, st);
  if (*e || !open || st->master_col > 0 || st->chunks.empty()) return e;
  double members = 0;
  for (const Chunk& ch : st->chunks) members += __builtin_popcount(ch.mask);
  members /= static_cast<double>(st->chunks.size());
  if (members >= 1.6) return e;
  Structure wide;
  const char* e2 = build_with(2, &wide);
  if (!*e2 && wide.tiles == 2) *st = std::move(wide);
  return e;
}

// ---- handle construction ---------------------------------------------------

static void release_staged(sparta_handle* h);
static void free_handle(sparta_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStream_t s = h->stream;
  if (s) {
    dev_free(h->d_segs, s); dev_free(h->d_srows, s); dev_free(h->d_chunks, s); dev_free(h->d_tables, s);
    dev_free(h->d_a, s); dev_free(h->d_items, s); dev_free(h->d_cta_ptr, s); dev_free(h->d_cta_items, s); dev_free(h->d_zero_jobs, s); dev_free(h->d_sync, s);
    dev_free(h->d_rowptr, s); dev_free(h->d_colind, s); dev_free(h->d_val, s); dev_free(h->d_row_order, s);
    dev_free(h->d_B, s); dev_free(h->d_B2, s); dev_free(h->d_C, s);
    for (auto& gp : h->gather_passes) { dev_free(gp.beg, s); dev_free(gp.end, s); dev_free(gp.order, s); }
    cudaStreamSynchronize(s);   // the blocks are back in the pool before the stream goes away
  }
  release_staged(h);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->up0) cudaEventDestroy(h->up0);
  if (h->up1) cudaEventDestroy(h->up1);
  if (s) cudaStreamDestroy(s);
  delete h;
}

// Host array -> device through a pinned block of the handle when it is large enough to matter (small
// pageable copies are embedded in the command stream and do not block either).
static cudaError_t staged_copy(sparta_handle* h, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return cudaSuccess;
  void* pin = bytes > (size_t{1} << 16) ? pinned_acquire(bytes) : nullptr;
  if (pin) {
    memcpy(pin, src, bytes);
    h->staged.push_back(pin);
    src = pin;
  }
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream);
}
// after a stream synchronisation: nothing reads the handle's staging blocks any more
static void release_staged(sparta_handle* h) {
  for (void* p : h->staged) pinned_release(p);
  h->staged.clear();
}
template <class T>
static cudaError_t upload_vec(const std::vector<T>& v, T** dptr, sparta_handle* h) {
  cudaError_t e = dev_alloc(dptr, v.size() * sizeof(T), h->stream);
  if (e != cudaSuccess) return e;
  return staged_copy(h, *dptr, v.data(), v.size() * sizeof(T));
}

// Device, stream and events of a new handle.  Returns SPARTA_OK or a failure code (handle freed).
static int open_handle(sparta_handle** out, const sparta_options& o, int* sms_out) {
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SPARTA_ERR_NO_DEVICE, "no CUDA device visible (libsparta_b200 has no CPU path)");
  }
  int dev = 0;
  if (o.device > 0) dev = o.device - 1; else CU_TRY(cudaGetDevice(&dev));
  CU_TRY(cudaSetDevice(dev));
  int major = 0, sms = 0;
  CU_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (major != 10)
    return fail(SPARTA_ERR_NO_DEVICE, "device is not compute capability 10.x (sm_100a kernels only)");
  if (o.precision < 0 || o.precision > 2) return fail(SPARTA_ERR_INVALID, "invalid precision");
  CU_TRY(pool_prepare(dev));
  sparta_handle* h = new sparta_handle();
  h->device = dev;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  if (e == cudaSuccess) e = cudaEventCreate(&h->up0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->up1);
  if (e != cudaSuccess) { free_handle(h); return fail_cuda(e, "stream / event creation"); }
  *sms_out = sms;
  *out = h;
  return SPARTA_OK;
}

// ---- sparse upload of a mostly-zero fp32 source -------------------------------------------------
struct SparseSource {
  std::vector<std::vector<int64_t>> idx;   // per scanning thread: element offsets of the nonzeros
  std::vector<std::vector<float>> val;
  int64_t total = 0;
  // or one caller-owned span; its offsets count from element `base` of the whole source (a shard's lists are
  // a slice of the matrix's lists: the scatter subtracts the shard's first element instead of the host)
  const int64_t* span_idx = nullptr;
  const float* span_val = nullptr;
  int64_t base = 0;
};

// fraction of nonzero elements in ~256 K sampled elements (4096 windows of 64)
static double sampled_density(const float* src, int64_t n) {
  const int64_t windows = 4096, len = 64;
  if (n < windows * len) return 1.0;
  int64_t nz = 0;
  uint64_t x = 0x9E3779B97F4A7C15ull;
  for (int64_t w = 0; w < windows; ++w) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    const int64_t at = static_cast<int64_t>(x % static_cast<uint64_t>(n - len));
    for (int64_t i = 0; i < len; ++i) nz += src[at + i] != 0.0f;
  }
  return static_cast<double>(nz) / static_cast<double>(windows * len);
}

// All host threads scan disjoint ranges; gives up (returns false) as soon as the source turns
// out denser than the sample suggested (12 bytes per nonzero against 4 per element).
static bool scan_nonzeros(const float* src, int64_t n, SparseSource* out) {
  const int T = host_thread_budget(32);
  out->idx.assign(T, {});
  out->val.assign(T, {});
  std::vector<char> gave_up(T, 0);
  const int64_t limit = n / 5 / T + 1024;   // per thread: at most 20 % nonzeros
  auto work = [&](int t) {
    const int64_t lo = n * t / T, hi = n * (t + 1) / T;
    std::vector<int64_t>& ix = out->idx[t];
    std::vector<float>& vl = out->val[t];
    ix.reserve(static_cast<size_t>((hi - lo) / 64 + 1024));
    vl.reserve(static_cast<size_t>((hi - lo) / 64 + 1024));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(src);
    int64_t i = lo;
    for (; i + 16 <= hi; i += 16) {
      uint32_t any = 0;
      for (int e = 0; e < 16; ++e) any |= w[i + e];
      if (!(any & 0x7FFFFFFFu)) continue;      // +0 / -0 only
      for (int e = 0; e < 16; ++e)
        if (w[i + e] & 0x7FFFFFFFu) { ix.push_back(i + e); vl.push_back(src[i + e]); }
      if (static_cast<int64_t>(ix.size()) > limit) { gave_up[t] = 1; return; }
    }
    for (; i < hi; ++i)
      if (w[i] & 0x7FFFFFFFu) { ix.push_back(i); vl.push_back(src[i]); }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  out->total = 0;
  for (int t = 0; t < T; ++t) {
    if (gave_up[t]) return false;
    out->total += static_cast<int64_t>(out->idx[t].size());
  }
  return true;
}

// The host tile scheduler (tens of milliseconds at 10^5 blocks) runs on its own thread WHILE the
// fp32 source crosses PCIe; nothing on the device waits for the CPU afterwards except the small
// schedule arrays.  No stream synchronisation happens here: the caller's next set_B / run /
// get_C are enqueued behind the upload on the handle's stream.
static int create_common(sparta_handle** out, BlockRows& br, const float* src_host,
                         int64_t src_elems, int64_t cols, const sparta_options& o,
                         int default_row_major, bool defer_sync = false, const SparseSource* pre = nullptr) {
  *out = nullptr;
  if (o.panel_stages < 2 || o.panel_stages > kMaxPanelStages)
    return fail(SPARTA_ERR_INVALID, "invalid panel_stages");
  const bool timing = getenv("SPARTA_TIMING") != nullptr;
  const auto tc0 = std::chrono::steady_clock::now();
  sparta_handle* h = nullptr;
  int sms = 0;
  const int rc = open_handle(&h, o, &sms);
  if (rc) return rc;
  h->sopt.precision = o.precision;
  h->sopt.seg_rows = o.seg_rows;
  h->sopt.acc_cols = o.acc_cols;
  h->sopt.pair = o.cta_pair != 1;
  h->panel_stages = o.panel_stages;
  h->a_ring_bytes = ring_bytes_for(o.panel_stages);
  {
    // The persistent grid must be co-resident (split plans synchronise the grid inside the
    // kernel): never launch more CTAs than the device holds at once at this footprint.
    int cap = 0;
    const cudaError_t ce = spmm_max_coresident_ctas(h->sopt.pair, o.precision == SPARTA_TF32, 0,
                                                    spmm_smem_bytes(h->panel_stages, h->a_ring_bytes), &cap);
    if (ce != cudaSuccess) { free_handle(h); return fail_cuda(ce, "cudaOccupancyMaxActiveClusters"); }
    if (cap <= 0) { free_handle(h); return fail(SPARTA_ERR_CUDA, "the device cannot hold a single CTA of the SpMM kernel"); }
    if (o.num_ctas > cap) {
      free_handle(h);
      return fail(SPARTA_ERR_INVALID, "num_ctas exceeds the CTAs the device can hold at the same time (" +
                                          std::to_string(cap) + "); the persistent grid must be co-resident");
    }
    h->sopt.num_ctas = o.num_ctas > 0 ? o.num_ctas : std::min(sms, cap);
  }
  h->sopt.sort_rows = o.row_order != 1;
  h->sopt.l2_slab_bytes = static_cast<int64_t>(o.l2_slab_mb) << 20;
  h->sopt.max_chain = o.max_chain;
  h->sopt.split = o.split_k;
  h->accumulate = o.accumulate ? 1 : 0;
  h->b_row_major = o.b_layout == SPARTA_LAYOUT_DEFAULT ? default_row_major : (o.b_layout == SPARTA_ROW_MAJOR);
  h->c_row_major = o.c_layout == SPARTA_LAYOUT_DEFAULT ? default_row_major : (o.c_layout == SPARTA_ROW_MAJOR);
  h->cols = cols;
  h->w = br.w;
  h->block_rows = br.count();

  const char* serr = "";
  // runs of short consecutive block-rows share one 16-row MMA segment (schedule.h)
  BlockRows fused;
  const BlockRows* view = &br;
  if (o.fuse_rows != 1 && fuse_short_block_rows(br, 16, &fused)) view = &fused;
  std::thread sched([&, view] {
    try {
      serr = build_structure_choosing_tiles(*view, o, &h->sopt, &h->st);
    } catch (...) {
      serr = "out of host memory while building the tile schedule";
    }
  });
  struct JoinGuard {   // an exception below must not reach the destructor of a joinable thread (std::terminate)
    std::thread& t;
    ~JoinGuard() { if (t.joinable()) t.join(); }
  } join_guard{sched};

#define H_TRY(call)                                                            \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      if (sched.joinable()) sched.join();                                      \
      cudaStreamSynchronize(h->stream);                                        \
      if (d_src) cudaFreeAsync(d_src, h->stream);                              \
      if (d_jobs) cudaFreeAsync(d_jobs, h->stream);                            \
      if (d_idx) cudaFreeAsync(d_idx, h->stream);                              \
      if (d_val) cudaFreeAsync(d_val, h->stream);                              \
      free_handle(h);                                                          \
      return fail_cuda(e_, #call);                                             \
    }                                                                          \
  } while (0)

  float* d_src = nullptr;
  PackJob* d_jobs = nullptr;
  int64_t* d_idx = nullptr;   // nonzeros-only upload
  float* d_val = nullptr;
  const auto tc_open = std::chrono::steady_clock::now();
  H_TRY(cudaEventRecord(h->up0, h->stream));
  bool sparse_upload = false;
  if (src_elems > 0) {
    // Stage the fp32 source on the device; it is repacked into MMA-ready images there.
    H_TRY(dev_alloc(&d_src, static_cast<size_t>(src_elems) * sizeof(float), h->stream));
    // The blocks SPARTA's clustering produces are mostly zeros inside (config #3: 3.5 M nonzeros
    // in 1.1 G stored elements), and PCIe is the slowest link of a one-shot call: when a sample
    // says the source is sparse, the host threads pick out the nonzeros (a read of the array at
    // memory speed), only those cross PCIe, and the dense image is rebuilt on the device.
    SparseSource sp;
    if (pre) {
      // the caller already holds the nonzeros (built from the CSR): only they cross PCIe
      sparse_upload = true;
      auto mark = [&](const char* what) {
        if (timing) fprintf(stderr, "  create mark %s %.1f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
      };
      mark("d_src allocated");
      H_TRY(cudaMemsetAsync(d_src, 0, static_cast<size_t>(src_elems) * sizeof(float), h->stream));
      mark("memset enqueued");
      if (pre->total > 0) {
        H_TRY(dev_alloc(&d_idx, static_cast<size_t>(pre->total) * sizeof(int64_t), h->stream));
        H_TRY(dev_alloc(&d_val, static_cast<size_t>(pre->total) * sizeof(float), h->stream));
        mark("idx/val allocated");
        int64_t at = 0;
        if (pre->span_idx) {
          H_TRY(cudaMemcpyAsync(d_idx, pre->span_idx, static_cast<size_t>(pre->total) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
          H_TRY(cudaMemcpyAsync(d_val, pre->span_val, static_cast<size_t>(pre->total) * sizeof(float), cudaMemcpyHostToDevice, h->stream));
          mark("copies enqueued");
        }
        for (size_t t = 0; t < pre->idx.size() && !pre->span_idx; ++t) {
          const size_t cnt = pre->idx[t].size();
          if (!cnt) continue;
          H_TRY(cudaMemcpyAsync(d_idx + at, pre->idx[t].data(), cnt * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
          H_TRY(cudaMemcpyAsync(d_val + at, pre->val[t].data(), cnt * sizeof(float), cudaMemcpyHostToDevice, h->stream));
          at += static_cast<int64_t>(cnt);
        }
        H_TRY(scatter_values(d_idx, d_val, pre->total, d_src - pre->base, h->stream));
        dev_free(d_idx, h->stream);
        dev_free(d_val, h->stream);
        d_idx = nullptr;
        d_val = nullptr;
      }
    } else if (src_elems >= (int64_t{1} << 22) && !getenv("SPARTA_DENSE_UPLOAD") &&
        sampled_density(src_host, src_elems) < 0.10) {
      // With a pinned source the copy engine and the host threads work at the same time: the head
      // of the array crosses PCIe as it is (the DMA runs at ~46 GB/s) while the threads scan the
      // tail (~64 GB/s on the 16-core bench host), split so that both finish together.
      int64_t head = 0;
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, src_host) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
        int pct = 50;   // share of the array the copy engine takes; SPARTA_UPLOAD_HEAD_PCT overrides (0..100)
        if (const char* e = getenv("SPARTA_UPLOAD_HEAD_PCT")) pct = std::max(0, std::min(100, atoi(e)));
        head = src_elems * pct / 100 / 64 * 64;
      }
      else
        cudaGetLastError();
      const int64_t tail = src_elems - head;
      H_TRY(cudaMemsetAsync(d_src + head, 0, static_cast<size_t>(tail) * sizeof(float), h->stream));
      if (head > 0)
        H_TRY(cudaMemcpyAsync(d_src, src_host, static_cast<size_t>(head) * sizeof(float), cudaMemcpyHostToDevice,
                              h->stream));
      if (scan_nonzeros(src_host + head, tail, &sp)) {
        sparse_upload = true;
        const int64_t nnz = sp.total;
        if (nnz > 0) {
          H_TRY(dev_alloc(&d_idx, static_cast<size_t>(nnz) * sizeof(int64_t), h->stream));
          H_TRY(dev_alloc(&d_val, static_cast<size_t>(nnz) * sizeof(float), h->stream));
          int64_t at = 0;
          for (size_t t = 0; t < sp.idx.size(); ++t) {
            const size_t cnt = sp.idx[t].size();
            if (!cnt) continue;
            H_TRY(cudaMemcpyAsync(d_idx + at, sp.idx[t].data(), cnt * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
            H_TRY(cudaMemcpyAsync(d_val + at, sp.val[t].data(), cnt * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            at += static_cast<int64_t>(cnt);
          }
          H_TRY(scatter_values(d_idx, d_val, nnz, d_src + head, h->stream));   // offsets are relative to the tail
          // pageable sources are staged before cudaMemcpyAsync returns; the vectors may go away
          dev_free(d_idx, h->stream);
          dev_free(d_val, h->stream);
          d_idx = nullptr;
          d_val = nullptr;
        }
      } else {
        // denser than the sample said: the tail goes up as it is
        H_TRY(cudaMemcpyAsync(d_src + head, src_host + head, static_cast<size_t>(tail) * sizeof(float),
                              cudaMemcpyHostToDevice, h->stream));
      }
    } else {
      H_TRY(cudaMemcpyAsync(d_src, src_host, static_cast<size_t>(src_elems) * sizeof(float),
                            cudaMemcpyHostToDevice, h->stream));
    }
  }
  const auto tc_enq = std::chrono::steady_clock::now();
  sched.join();
  const auto tc1 = std::chrono::steady_clock::now();
  if (*serr || static_cast<int64_t>(h->st.max_chunk_bytes) > h->a_ring_bytes) {
    cudaStreamSynchronize(h->stream);
    if (d_src) cudaFreeAsync(d_src, h->stream);
    free_handle(h);
    return fail(SPARTA_ERR_INVALID, *serr ? serr
                : "a chunk's A images exceed the shared-memory ring; lower acc_cols or panel_stages");
  }
  h->rows = h->st.rows;
  pick_pipeline(h->st.max_chunk_bytes, o, h->st.tiles, &h->panel_stages, &h->a_ring_bytes, &h->a_slot_bytes, &h->producers);
  H_TRY(upload_vec(h->st.segs, &h->d_segs, h));
  H_TRY(upload_vec(h->st.srows, &h->d_srows, h));
  H_TRY(upload_vec(h->st.chunks, &h->d_chunks, h));
  H_TRY(upload_vec(h->st.tables, &h->d_tables, h));
  H_TRY(dev_alloc(&h->d_a, h->st.a_bytes, h->stream));
  if (h->st.sparse_images && h->st.a_bytes) H_TRY(cudaMemsetAsync(h->d_a, 0, h->st.a_bytes, h->stream));
  if (h->st.n_jobs > 0) {
    H_TRY(dev_alloc(&d_jobs, static_cast<size_t>(h->st.n_jobs) * sizeof(PackJob), h->stream));
    size_t at = 0;
    for (const auto& part : h->st.job_parts) {
      if (part.empty()) continue;
      H_TRY(staged_copy(h, d_jobs + at, part.data(), part.size() * sizeof(PackJob)));
      at += part.size();
    }
    H_TRY(pack_a_images(d_src, d_jobs, h->st.n_jobs, h->d_a, o.precision, h->stream, static_cast<int64_t>(h->st.a_bytes)));
  }
  dev_free(d_src, h->stream);
  dev_free(d_jobs, h->stream);
  std::vector<std::vector<PackJob>>().swap(h->st.job_parts);   // pageable copies are staged before cudaMemcpyAsync returns
  H_TRY(cudaEventRecord(h->up1, h->stream));
  // The public create returns only when the caller's arrays are no longer being read.
  if (!defer_sync) {
    H_TRY(cudaStreamSynchronize(h->stream));
    release_staged(h);
  }
  if (timing)
    fprintf(stderr, "sparta create: [open %.1f, A upload enqueued %.1f, wait for the schedule %.1f] host schedule + enqueue of the A upload (%s) %.1f ms, enqueue of the rest %.1f ms\n",
            std::chrono::duration<double, std::milli>(tc_open - tc0).count(),
            std::chrono::duration<double, std::milli>(tc_enq - tc_open).count(),
            std::chrono::duration<double, std::milli>(tc1 - tc_enq).count(),
            sparse_upload ? "nonzeros only" : "dense",
            std::chrono::duration<double, std::milli>(tc1 - tc0).count(),
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc1).count());
#undef H_TRY
  *out = h;
  return SPARTA_OK;
}

// ---- extern "C" ------------------------------------------------------------

extern "C" {

const char* sparta_last_error(void) { return g_last_error.c_str(); }
int sparta_abi_version(void) { return SPARTA_ABI_VERSION; }

int sparta_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

// Uploads the gather rows of a hybrid handle (values rounded to the operand precision on the device).
// On failure the handle is freed.
static int attach_gather(sparta_handle* h, const GatherPart& gp, int precision) {
  cudaError_t e = upload_vec(gp.colind, &h->d_colind, h);
  if (e == cudaSuccess) e = upload_vec(gp.val, &h->d_val, h);
  h->gather_passes.assign(gp.passes.size(), sparta_handle::GatherPassDev());
  for (size_t pi = 0; pi < gp.passes.size() && e == cudaSuccess; ++pi) {
    sparta_handle::GatherPassDev& d = h->gather_passes[pi];
    e = upload_vec(gp.passes[pi].beg, &d.beg, h);
    if (e == cudaSuccess) e = upload_vec(gp.passes[pi].end, &d.end, h);
    if (e == cudaSuccess) e = upload_vec(gp.passes[pi].order, &d.order, h);
    d.work = static_cast<int64_t>(gp.passes[pi].order.size());
    d.heavy = gp.passes[pi].heavy;
  }
  if (e == cudaSuccess) e = round_values(h->d_val, static_cast<int64_t>(gp.val.size()), precision, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);   // the host vectors go away
  if (e != cudaSuccess) { free_handle(h); return fail_cuda(e, "upload of the gather rows"); }
  release_staged(h);
  h->gather_rows = gp.rows;
  h->gather_work = gp.passes.empty() ? 0 : static_cast<int64_t>(gp.passes[0].order.size());
  h->nnz = static_cast<int64_t>(gp.val.size());
  h->st.nztot += gp.nztot;
  h->st.n_blocks += gp.blocks;
  return SPARTA_OK;
}

// A straight from the flat CSR and the row grouping: VBR::fill_from_CSR_inplace (vbr.cpp:135-237) and
// the upload half of the multiply routines in one step.  The host only builds the INDEX arrays
// (row_part / nzcount / jab, bit-identical to the reference's) and the element offset every nonzero
// would have in mab; the dense blocks are rebuilt on the device from those (offset, value) pairs and
// packed there.  BASELINE config #3: 3.5 M nonzeros = 42 MB over PCIe instead of a 4.45 GB mab that is
// 99.7 % zeros.  The gather rows of a hybrid handle are CSR rows as they are.
static int vbr_create_from_csr_impl(sparta_handle** out, int64_t rows, int64_t cols, const int64_t* rowptr,
                                    const int64_t* colind, const float* val, const int64_t* grouping,
                                    int64_t block_col_size, int64_t row_block_size, int32_t force_fixed_size,
                                    const sparta_options* opt, bool defer_sync, HostVBR* index_out) {
  if (!out) return fail(SPARTA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (rows <= 0 || cols <= 0 || !rowptr || !grouping || (rowptr[rows] && !colind) || block_col_size <= 0)
    return fail(SPARTA_ERR_INVALID, "invalid CSR input");
  sparta_options o;
  resolve_options(opt, &o);
  const int threads = host_thread_budget(32);
  HostVBRSparse hv;
  // the nonzero lists are written by the filling threads straight into page-locked memory: they cross PCIe by DMA
  hv.nz_off.acquire_fn = pinned_or_malloc; hv.nz_off.release_fn = pinned_or_free;
  hv.nz_val.acquire_fn = pinned_or_malloc; hv.nz_val.release_fn = pinned_or_free;
  const char* e = host_vbr_fill_sparse(rows, cols, rowptr, colind, val, val == nullptr, grouping, block_col_size,
                                       row_block_size, force_fixed_size != 0, threads, &hv);
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  const HostVBR& ix = hv.index;
  const int64_t lo = o.block_row_begin;
  const int64_t hi = range_end(o, ix.block_rows);
  BlockRows br;
  int64_t src_lo = 0, src_hi = 0;
  e = blockrows_from_vbr(ix.block_rows, block_col_size, ix.row_part.data(), ix.nzcount.data(), ix.jab.data(), lo, hi,
                         &br, &src_lo, &src_hi);
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  const int64_t shard_rows = ix.row_part[hi] - ix.row_part[lo];
  // gather rows: the CSR rows of the short block-rows, as they are
  const int gather_h = o.gather_max_height < 0 ? 0 : (o.gather_max_height == 0 ? 7 : o.gather_max_height);
  std::vector<char> is_short(static_cast<size_t>(hi - lo), 0);
  GatherPart gp;
  bool hybrid = false;
  if (gather_h > 0 && shard_rows <= INT32_MAX && ix.cols <= INT32_MAX) {
    gp.rowptr.assign(static_cast<size_t>(shard_rows) + 1, 0);
    std::vector<int32_t> live;
    for (int64_t b = lo; b < hi; ++b) {
      const int64_t H = ix.row_part[b + 1] - ix.row_part[b];
      if (H <= 0 || H > gather_h || ix.nzcount[b] == 0) continue;
      is_short[static_cast<size_t>(b - lo)] = 1;
      hybrid = true;
      gp.rows += H;
      gp.blocks += ix.nzcount[b];
      gp.nztot += ix.nzcount[b] * H * block_col_size;
    }
    if (hybrid) {
      for (int64_t b = lo; b < hi; ++b) {
        for (int64_t r = ix.row_part[b]; r < ix.row_part[b + 1]; ++r) {
          int64_t cnt = 0;
          if (is_short[static_cast<size_t>(b - lo)] && r < rows) cnt = rowptr[hv.perm[r] + 1] - rowptr[hv.perm[r]];
          gp.rowptr[static_cast<size_t>(r - ix.row_part[lo]) + 1] = cnt;
        }
      }
      for (int64_t r = 0; r < shard_rows; ++r) gp.rowptr[r + 1] += gp.rowptr[r];
      gp.colind.resize(static_cast<size_t>(gp.rowptr[shard_rows]));
      gp.val.resize(gp.colind.size());
      for (int64_t b = lo; b < hi; ++b) {
        if (!is_short[static_cast<size_t>(b - lo)]) continue;
        for (int64_t r = ix.row_part[b]; r < ix.row_part[b + 1] && r < rows; ++r) {
          const int64_t i = hv.perm[r], local = r - ix.row_part[lo];
          int64_t at = gp.rowptr[local];
          for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p, ++at) {
            gp.colind[at] = static_cast<int32_t>(colind[p]);
            gp.val[at] = val ? val[p] : 1.0f;
          }
          if (rowptr[i + 1] > rowptr[i]) live.push_back(static_cast<int32_t>(local));
        }
      }
      build_gather_passes(&gp, live, shard_rows, ix.cols, prec_esize(o.precision), o.gather_passes);
    }
  }
  // the tile schedule's view: the tall block-rows; its nonzeros: (offset relative to the shard, value)
  BlockRows tall;
  if (hybrid) {
    tall.w = br.w;
    tall.ptr.push_back(0);
    for (int64_t b = 0; b < br.count(); ++b) {
      if (is_short[static_cast<size_t>(b)]) continue;
      tall.row0.push_back(br.row0[b]); tall.height.push_back(br.height[b]); tall.rs.push_back(br.rs[b]); tall.ks.push_back(br.ks[b]);
      for (int64_t q = br.ptr[b]; q < br.ptr[b + 1]; ++q) { tall.col.push_back(br.col[q]); tall.src.push_back(br.src[q]); }
      tall.ptr.push_back(static_cast<int64_t>(tall.col.size()));
    }
  }
  SparseSource pre;
  if (!hybrid) {
    // the whole matrix or a shard of it without gather rows: the lists (a slice of them) as they are
    pre.span_idx = hv.nz_off.p + hv.nz_ptr[lo];
    pre.span_val = hv.nz_val.p + hv.nz_ptr[lo];
    pre.total = hv.nz_ptr[hi] - hv.nz_ptr[lo];
    pre.base = src_lo;
  } else {
    pre.idx.assign(1, {});
    pre.val.assign(1, {});
    pre.idx[0].reserve(static_cast<size_t>(hv.nz_ptr[hi] - hv.nz_ptr[lo]));
    pre.val[0].reserve(static_cast<size_t>(hv.nz_ptr[hi] - hv.nz_ptr[lo]));
    for (int64_t b = lo; b < hi; ++b) {
      if (is_short[static_cast<size_t>(b - lo)]) continue;
      for (int64_t q = hv.nz_ptr[b]; q < hv.nz_ptr[b + 1]; ++q) {
        pre.idx[0].push_back(hv.nz_off[q] - src_lo);
        pre.val[0].push_back(hv.nz_val[q]);
      }
    }
    pre.total = static_cast<int64_t>(pre.idx[0].size());
  }
  // The one-shot callers defer the synchronisation: pageable lists are staged by the driver before the copy
  // call returns, page-locked ones (the spans) are handed to the handle, which releases them after its next
  // synchronisation.
  const int rc = create_common(out, hybrid ? tall : br, nullptr, src_hi - src_lo, ix.cols, o, 0, defer_sync, &pre);
  if (rc != SPARTA_OK) return rc;
  sparta_handle* h = *out;
  if (defer_sync && pre.span_idx) {
    if (hv.nz_off.tag) { h->staged.push_back(hv.nz_off.p); hv.nz_off.p = nullptr; }
    if (hv.nz_val.tag) { h->staged.push_back(hv.nz_val.p); hv.nz_val.p = nullptr; }
  }
  h->rows = shard_rows;
  h->st.rows = shard_rows;
  h->block_rows = hi - lo;
  if (hybrid) {
    const int grc = attach_gather(h, gp, o.precision);
    if (grc != SPARTA_OK) { *out = nullptr; return grc; }
  }
  if (index_out) *index_out = std::move(hv.index);
  return SPARTA_OK;
}

static int vbr_create_impl(sparta_handle** out, int64_t rows, int64_t cols, int64_t block_rows,
                           int64_t block_col_size, const int64_t* row_part, const int64_t* nzcount,
                           const int64_t* jab, const float* mab, const sparta_options* opt,
                           bool defer_sync) {
  if (!out) return fail(SPARTA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (rows < 0 || cols <= 0 || block_rows < 0 || block_col_size <= 0 || !row_part || (block_rows && !nzcount))
    return fail(SPARTA_ERR_INVALID, "invalid VBR dimensions or NULL index arrays");
  if (block_rows > 0 && row_part[block_rows] != rows)
    return fail(SPARTA_ERR_INVALID, "row_part[block_rows] must equal rows (VBR::partition_check)");
  sparta_options o;
  resolve_options(opt, &o);
  const int64_t lo = o.block_row_begin;
  const int64_t hi = range_end(o, block_rows);
  BlockRows br;
  int64_t src_lo = 0, src_hi = 0;
  const char* e = blockrows_from_vbr(block_rows, block_col_size, row_part, nzcount, jab, lo, hi, &br, &src_lo, &src_hi);
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  const int64_t bc = (cols - 1) / block_col_size + 1;
  for (int64_t jb : br.col)
    if (jb >= bc) return fail(SPARTA_ERR_INVALID, "jab entry beyond the last column block");
  if (src_hi > src_lo && !mab) return fail(SPARTA_ERR_INVALID, "mab is NULL");
  const int gather_h = o.gather_max_height < 0 ? 0 : (o.gather_max_height == 0 ? 7 : o.gather_max_height);
  BlockRows tall;
  GatherPart gp;
  const bool hybrid = mab && split_short_block_rows(br, gather_h, mab + src_lo, cols, prec_esize(o.precision), o.gather_passes, &tall, &gp);
  const int64_t shard_rows = row_part[hi] - row_part[lo];
  const int rc = create_common(out, hybrid ? tall : br, mab ? mab + src_lo : nullptr, src_hi - src_lo, cols, o, 0,
                               defer_sync && !hybrid);
  if (rc != SPARTA_OK) return rc;
  sparta_handle* h = *out;
  h->rows = shard_rows;           // the tile schedule may not reach the last rows any more
  h->st.rows = shard_rows;
  h->block_rows = hi - lo;
  if (hybrid) {
    const int grc = attach_gather(h, gp, o.precision);
    if (grc != SPARTA_OK) { *out = nullptr; return grc; }
  }
  return SPARTA_OK;
}

int sparta_vbr_create(sparta_handle** out, int64_t rows, int64_t cols, int64_t block_rows,
                      int64_t block_col_size, const int64_t* row_part, const int64_t* nzcount,
                      const int64_t* jab, const float* mab, const sparta_options* opt) {
  return guarded([&] { return vbr_create_impl(out, rows, cols, block_rows, block_col_size, row_part, nzcount, jab, mab, opt, false); });
}

int sparta_vbr_create_from_csr(sparta_handle** out, int64_t rows, int64_t cols, const int64_t* rowptr,
                               const int64_t* colind, const float* val, const int64_t* grouping,
                               int64_t block_col_size, int64_t row_block_size, int32_t force_fixed_size,
                               const sparta_options* opt, int64_t* dims) {
  HostVBR ix;
  const int rc = guarded([&] {
    return vbr_create_from_csr_impl(out, rows, cols, rowptr, colind, val, grouping, block_col_size, row_block_size,
                                    force_fixed_size, opt, false, &ix);
  });
  if (rc == SPARTA_OK && dims) {
    dims[0] = ix.rows; dims[1] = ix.cols; dims[2] = ix.block_rows; dims[3] = ix.block_cols;
    dims[4] = ix.block_col_size; dims[5] = ix.nztot;
  }
  return rc;
}

static int vbr_create_ba_impl(sparta_handle** out, int64_t rows, int64_t cols, int64_t block_rows,
                              int64_t block_col_size, const int64_t* row_part, const int64_t* nzcount,
                              const int64_t* jab, const float* mab, const sparta_options* opt,
                              bool defer_sync) {
  if (!out) return fail(SPARTA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (rows <= 0 || cols <= 0 || block_rows < 0 || block_col_size <= 0 || !row_part || (block_rows && !nzcount))
    return fail(SPARTA_ERR_INVALID, "invalid VBR dimensions or NULL index arrays");
  if (block_rows > 0 && row_part[block_rows] != rows)
    return fail(SPARTA_ERR_INVALID, "row_part[block_rows] must equal rows (VBR::partition_check)");
  sparta_options o;
  resolve_options(opt, &o);
  const int64_t bc = (cols - 1) / block_col_size + 1;
  const int64_t lo = o.block_row_begin;
  const int64_t hi = range_end(o, bc);
  BlockRows br;
  int64_t src_hi = 0;
  const char* e = blockrows_from_vbr_transposed(cols, block_rows, block_col_size, row_part, nzcount, jab, lo, hi,
                                                &br, &src_hi);
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  if (src_hi > 0 && !mab) return fail(SPARTA_ERR_INVALID, "mab is NULL");
  // the operand's k runs over A's rows; B (n x rows) and C (n x cols) column-major are the
  // row-major [rows][n] and [cols][n] operands of C^T = A^T * B^T
  const int rc = create_common(out, br, mab, src_hi, rows, o, 1, defer_sync);
  if (rc == SPARTA_OK) {   // FLOP accounting on the reference's nztot (full-width last column block)
    int64_t blocks = 0;
    for (int64_t t = 0; t < static_cast<int64_t>(br.blk_kw.size()); ++t) blocks += br.blk_kw[t];
    (*out)->st.nztot = blocks * block_col_size;
  }
  return rc;
}

int sparta_vbr_create_BA(sparta_handle** out, int64_t rows, int64_t cols, int64_t block_rows,
                         int64_t block_col_size, const int64_t* row_part, const int64_t* nzcount,
                         const int64_t* jab, const float* mab, const sparta_options* opt) {
  return guarded([&] { return vbr_create_ba_impl(out, rows, cols, block_rows, block_col_size, row_part, nzcount, jab, mab, opt, false); });
}

static int bellpack_create_impl(sparta_handle** out, int64_t rows, int64_t cols, int64_t ell_blocksize,
                                int64_t ellColInd_rows, int64_t ellColInd_cols,
                                const int64_t* ellColInd, const float* ellValues,
                                const sparta_options* opt, bool defer_sync) {
  if (!out) return fail(SPARTA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (ell_blocksize <= 0 || rows < 0 || cols <= 0 || ellColInd_rows < 0 || ellColInd_cols < 0)
    return fail(SPARTA_ERR_INVALID, "invalid Blocked-ELL dimensions");
  // same shape rule as prepare_cusparse_BLOCKEDELLPACK (cuda_utilities.cpp:1664-1670)
  if (rows % ell_blocksize || cols % ell_blocksize || ellColInd_rows != rows / ell_blocksize)
    return fail(SPARTA_ERR_INVALID, "rows and cols must be multiples of ell_blocksize");
  if (ellColInd_rows * ellColInd_cols > 0 && (!ellColInd || !ellValues))
    return fail(SPARTA_ERR_INVALID, "NULL Blocked-ELL arrays");
  sparta_options o;
  resolve_options(opt, &o);
  const int64_t lo = o.block_row_begin;
  const int64_t hi = range_end(o, ellColInd_rows);
  BlockRows br;
  int64_t src_lo = 0, src_hi = 0;
  const char* e = blockrows_from_bell(ell_blocksize, ellColInd_rows, ellColInd_cols, ellColInd, lo, hi, &br, &src_lo, &src_hi);
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  for (int64_t jb : br.col)
    if (jb >= cols / ell_blocksize) return fail(SPARTA_ERR_INVALID, "ellColInd entry beyond the last column block");
  return create_common(out, br, ellValues ? ellValues + src_lo : nullptr, src_hi - src_lo, cols, o, 1, defer_sync);
}

int sparta_bellpack_create(sparta_handle** out, int64_t rows, int64_t cols, int64_t ell_blocksize,
                           int64_t ellColInd_rows, int64_t ellColInd_cols,
                           const int64_t* ellColInd, const float* ellValues,
                           const sparta_options* opt) {
  return guarded([&] {
    return bellpack_create_impl(out, rows, cols, ell_blocksize, ellColInd_rows, ellColInd_cols, ellColInd, ellValues, opt, false);
  });
}

// CSR handles: rows of the shard [row_begin, row_end) (block_row_begin / block_row_end of the
// options count ROWS here).  Index arrays are narrowed to int32 columns on the host while the
// values cross PCIe; the row order (descending nnz, stable) is the kernel's work list.
static int csr_create_impl(sparta_handle** out, int64_t rows, int64_t cols, const int64_t* rowptr,
                           const int64_t* colind, const float* val, const sparta_options* opt,
                           bool defer_sync) {
  if (!out) return fail(SPARTA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (rows < 0 || cols <= 0 || !rowptr || (rows && rowptr[rows] > rowptr[0] && !colind))
    return fail(SPARTA_ERR_INVALID, "invalid CSR dimensions or NULL index arrays");
  if (cols > INT32_MAX || rows > INT32_MAX) return fail(SPARTA_ERR_INVALID, "CSR dimensions exceed int32");
  sparta_options o;
  resolve_options(opt, &o);
  const int64_t lo = o.block_row_begin;
  const int64_t hi = range_end(o, rows);
  if (lo < 0 || hi > rows || lo > hi) return fail(SPARTA_ERR_INVALID, "row range out of bounds");
  for (int64_t i = lo; i < hi; ++i)
    if (rowptr[i + 1] < rowptr[i]) return fail(SPARTA_ERR_INVALID, "rowptr must be non-decreasing");
  const int64_t p0 = rowptr[lo], nnz = rowptr[hi] - rowptr[lo], nrows = hi - lo;
  sparta_handle* h = nullptr;
  int sms = 0;
  const int rc = open_handle(&h, o, &sms);
  if (rc) return rc;
  h->kind = 1;
  h->sopt.precision = o.precision;
  h->accumulate = o.accumulate ? 1 : 0;
  h->b_row_major = o.b_layout == SPARTA_LAYOUT_DEFAULT ? 1 : (o.b_layout == SPARTA_ROW_MAJOR);
  h->c_row_major = o.c_layout == SPARTA_LAYOUT_DEFAULT ? 1 : (o.c_layout == SPARTA_ROW_MAJOR);
  h->cols = cols;
  h->rows = nrows;
  h->nnz = nnz;
  h->block_rows = nrows;
  h->w = 1;
  h->st.rows = nrows;
  h->st.nztot = nnz;
  h->st.n_blocks = nnz;
#define H_TRY(call)                                                            \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) { free_handle(h); return fail_cuda(e_, #call); }    \
  } while (0)
  H_TRY(cudaEventRecord(h->up0, h->stream));
  H_TRY(dev_alloc(&h->d_val, static_cast<size_t>(nnz) * sizeof(float), h->stream));
  std::vector<float> ones;
  if (nnz > 0) {
    if (val) {
      H_TRY(cudaMemcpyAsync(h->d_val, val + p0, static_cast<size_t>(nnz) * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    } else {   // pattern-only: every entry is 1 (csr.cpp:59)
      ones.assign(static_cast<size_t>(nnz), 1.f);
      H_TRY(cudaMemcpyAsync(h->d_val, ones.data(), static_cast<size_t>(nnz) * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    }
  }
  std::vector<int32_t> col32(static_cast<size_t>(nnz));
  for (int64_t q = 0; q < nnz; ++q) {
    const int64_t c = colind[p0 + q];
    if (c < 0 || c >= cols) { free_handle(h); return fail(SPARTA_ERR_INVALID, "column index out of range"); }
    col32[static_cast<size_t>(q)] = static_cast<int32_t>(c);
  }
  std::vector<int64_t> ptr(static_cast<size_t>(nrows) + 1);
  for (int64_t i = 0; i <= nrows; ++i) ptr[static_cast<size_t>(i)] = rowptr[lo + i] - p0;
  std::vector<int32_t> order(static_cast<size_t>(nrows));
  for (int64_t i = 0; i < nrows; ++i) order[static_cast<size_t>(i)] = static_cast<int32_t>(i);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    return ptr[a + 1] - ptr[a] > ptr[b + 1] - ptr[b];
  });
  h->heavy_rows = 0;
  while (h->heavy_rows < nrows && ptr[order[h->heavy_rows] + 1] - ptr[order[h->heavy_rows]] > kCsrHeavyNnz)
    ++h->heavy_rows;
  H_TRY(upload_vec(col32, &h->d_colind, h));
  H_TRY(upload_vec(ptr, &h->d_rowptr, h));
  H_TRY(upload_vec(order, &h->d_row_order, h));
  H_TRY(cudaEventRecord(h->up1, h->stream));
  if (!defer_sync || !val) {
    H_TRY(cudaStreamSynchronize(h->stream));
    release_staged(h);
  }
#undef H_TRY
  *out = h;
  return SPARTA_OK;
}

int sparta_csr_create(sparta_handle** out, int64_t rows, int64_t cols, const int64_t* rowptr,
                      const int64_t* colind, const float* val, const sparta_options* opt) {
  return guarded([&] { return csr_create_impl(out, rows, cols, rowptr, colind, val, opt, false); });
}

static int set_b_impl(sparta_handle* h, const float* B, int64_t ld, int64_t n, int on_device,
                      bool defer_sync) {
  if (!h || !B) return fail(SPARTA_ERR_INVALID, "NULL handle or B");
  if (n <= 0) return fail(SPARTA_ERR_INVALID, "n must be positive");
  const int64_t min_ld = h->b_row_major ? n : h->cols;
  if (ld < min_ld) return fail(SPARTA_ERR_INVALID, "leading dimension of B too small");
  if (h->kind == 1 && n > (1 << 30)) return fail(SPARTA_ERR_INVALID, "n too large");
  CU_TRY(cudaSetDevice(h->device));
  h->chain_ok = false;   // B, the work assignment and C's buffer change below
  CU_TRY(cudaEventRecord(h->up0, h->stream));
  const bool timing = getenv("SPARTA_TIMING") != nullptr;
  const auto ts0 = std::chrono::steady_clock::now();
  auto ts_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts0).count(); };
  double ts_conv = 0, ts_assign = 0;
  const int esize = (h->kind == 1 && h->sopt.precision == PREC_TF32) ? 4 : prec_esize(h->sopt.precision);
  // block kernel: [n][ldk] k-contiguous;  CSR kernel: [cols][ldn] n-contiguous, ldn = n rounded to 8
  const int64_t ldk = h->kind == 1 ? (n + 7) / 8 * 8 : (h->cols + 63) / 64 * 64;
  const size_t b_bytes = static_cast<size_t>(h->kind == 1 ? h->cols : n) * ldk * esize;
  if (b_bytes > h->b_cap) {
    dev_free(h->d_B, h->stream);
    h->b_cap = 0;
    CU_TRY(dev_alloc(&h->d_B, b_bytes, h->stream));
    h->b_cap = b_bytes;
  }
  h->ldk = ldk;
  // fp32 staging copy (host sources only)
  const float* src = B;
  float* d_stage = nullptr;
  int64_t src_ld = ld;
  if (!on_device) {
    const int64_t lines = h->b_row_major ? h->cols : n;     // number of ld-strided lines
    const int64_t width = h->b_row_major ? n : h->cols;     // contiguous elements per line
    CU_TRY(dev_alloc(&d_stage, static_cast<size_t>(lines) * width * sizeof(float), h->stream));
    cudaError_t e = cudaMemcpy2DAsync(d_stage, width * sizeof(float), B, ld * sizeof(float),
                                      width * sizeof(float), lines, cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) { dev_free(d_stage, h->stream); return fail_cuda(e, "B upload"); }
    src = d_stage;
    src_ld = width;
  }
  cudaError_t e;
  if (h->kind == 1) {
    // the CSR kernel wants rows of B contiguous: same converters with the roles of the two
    // dimensions exchanged (a row-major source needs no transpose here), zero-filled padding
    if (ldk != n) e = cudaMemsetAsync(h->d_B, 0, b_bytes, h->stream); else e = cudaSuccess;
    if (e == cudaSuccess)
      e = convert_b(src, src_ld, !h->b_row_major, h->d_B, ldk, n, h->cols, h->sopt.precision,
                    h->stream, /*keep_fp32=*/h->sopt.precision == PREC_TF32);
  } else {
    e = convert_b(src, src_ld, h->b_row_major, h->d_B, ldk, h->cols, n, h->sopt.precision, h->stream, false);
    if (e == cudaSuccess && h->gather_work > 0) {
      // the gather rows read ROWS of B: a second copy [cols][ldn] in the operand precision (tf32:
      // 4-byte values rounded like the tensor-core operand)
      const int64_t ldn = (n + 7) / 8 * 8;
      const size_t b2 = static_cast<size_t>(h->cols) * ldn * prec_esize(h->sopt.precision);
      if (b2 > h->b2_cap) {
        dev_free(h->d_B2, h->stream);
        h->b2_cap = 0;
        e = dev_alloc(&h->d_B2, b2, h->stream);
        if (e == cudaSuccess) h->b2_cap = b2;
      }
      if (e == cudaSuccess && ldn != n) e = cudaMemsetAsync(h->d_B2, 0, b2, h->stream);
      if (e == cudaSuccess)
        e = convert_b(src, src_ld, !h->b_row_major, h->d_B2, ldn, n, h->cols, h->sopt.precision, h->stream, false);
      h->ldn2 = ldn;
    }
  }
  dev_free(d_stage, h->stream);
  if (e != cudaSuccess) return fail_cuda(e, "B conversion");
  ts_conv = ts_ms();

  if (n != h->n) {
    if (h->kind == 0) {
      const char* serr = build_assignment(h->st, h->sopt, n, h->cols, &h->as);
      if (*serr) return fail(SPARTA_ERR_INVALID, serr);
      ts_assign = ts_ms();
      dev_free(h->d_items, h->stream); dev_free(h->d_cta_ptr, h->stream); dev_free(h->d_cta_items, h->stream);
      dev_free(h->d_zero_jobs, h->stream);
      h->d_zero_jobs = nullptr;
      if (!h->as.zero_jobs.empty()) CU_TRY(upload_vec(h->as.zero_jobs, &h->d_zero_jobs, h));
      CU_TRY(upload_vec(h->as.items, &h->d_items, h));
      CU_TRY(upload_vec(h->as.cta_ptr, &h->d_cta_ptr, h));
      CU_TRY(upload_vec(h->as.cta_items, &h->d_cta_items, h));
    }
    // C: column-major ld padded to 4 rows so the epilogue can use 16-byte stores
    h->ldc = h->c_row_major ? n : (h->rows + 3) / 4 * 4;
    const size_t c_elems = h->c_row_major ? static_cast<size_t>(h->rows) * n : static_cast<size_t>(h->ldc) * n;
    const size_t c_bytes = std::max<size_t>(c_elems, 4) * sizeof(float);
    if (c_bytes > h->c_cap) {
      dev_free(h->d_C, h->stream);
      h->c_cap = 0;
      CU_TRY(dev_alloc(&h->d_C, c_bytes, h->stream));
      h->c_cap = c_bytes;
    }
    CU_TRY(cudaMemsetAsync(h->d_C, 0, c_bytes, h->stream));
    h->n = n;
  }
  CU_TRY(cudaEventRecord(h->up1, h->stream));
  if (timing)
    fprintf(stderr, "sparta set_B: upload / conversion enqueued %.1f ms, work assignment %.1f, its upload + C %.1f\n", ts_conv,
            ts_assign > 0 ? ts_assign - ts_conv : 0.0, ts_ms() - (ts_assign > 0 ? ts_assign : ts_conv));
  if (!defer_sync) {
    CU_TRY(cudaStreamSynchronize(h->stream));
    release_staged(h);
  }
  return SPARTA_OK;
}

int sparta_set_B(sparta_handle* h, const float* B, int64_t ld, int64_t n, int on_device) {
  return guarded([&] { return set_b_impl(h, B, ld, n, on_device, false); });
}

static int copy_c(sparta_handle* h, float* C, int64_t ld, int on_device, bool to_handle) {
  if (!h || !C) return fail(SPARTA_ERR_INVALID, "NULL handle or C");
  if (h->n == 0) return fail(SPARTA_ERR_STATE, "set_B must be called first");
  const int64_t rows = h->rows;
  const int64_t lines = h->c_row_major ? rows : h->n;
  const int64_t width = h->c_row_major ? h->n : rows;
  if (ld < width) return fail(SPARTA_ERR_INVALID, "leading dimension of C too small");
  if (lines == 0 || width == 0) return SPARTA_OK;
  CU_TRY(cudaSetDevice(h->device));
  h->chain_ok = false;   // something other than the handle's own multiply is about to sit on the stream
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice
                              : (to_handle ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost);
  if (to_handle)
    CU_TRY(cudaMemcpy2DAsync(h->d_C, h->ldc * sizeof(float), C, ld * sizeof(float), width * sizeof(float), lines, kind, h->stream));
  else
    CU_TRY(cudaMemcpy2DAsync(C, ld * sizeof(float), h->d_C, h->ldc * sizeof(float), width * sizeof(float), lines, kind, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  release_staged(h);
  return SPARTA_OK;
}

int sparta_set_C(sparta_handle* h, const float* C, int64_t ld, int on_device) {
  // with accumulate = 0 the multiply overwrites C, except the rows of empty block-rows, which
  // rely on C's initial zero fill (build_assignment): an initial C only makes sense for beta = 1
  if (h && !h->accumulate) return fail(SPARTA_ERR_STATE, "sparta_set_C needs a handle created with accumulate = 1");
  return copy_c(h, const_cast<float*>(C), ld, on_device, true);
}
int sparta_get_C(sparta_handle* h, float* C, int64_t ld, int on_device) {
  return copy_c(h, C, ld, on_device, false);
}

// C with its rows moved to row_map[r]: the scatter runs on the device (HBM-bound, 2 x C bytes).
int sparta_get_C_permuted(sparta_handle* h, float* C, int64_t ld, const int64_t* row_map, int64_t out_rows,
                          int on_device) {
  if (!h || !C || !row_map) return fail(SPARTA_ERR_INVALID, "NULL handle, C or row_map");
  if (h->n == 0) return fail(SPARTA_ERR_STATE, "set_B must be called first");
  const int64_t rows = h->rows, n = h->n;
  if (out_rows < rows) return fail(SPARTA_ERR_INVALID, "out_rows smaller than the handle's rows");
  const int64_t width = h->c_row_major ? n : out_rows;
  if (ld < width) return fail(SPARTA_ERR_INVALID, "leading dimension of C too small");
  for (int64_t r = 0; r < rows; ++r)
    if (row_map[r] < 0 || row_map[r] >= out_rows) return fail(SPARTA_ERR_INVALID, "row_map entry out of range");
  if (rows == 0 || n == 0) return SPARTA_OK;
  CU_TRY(cudaSetDevice(h->device));
  h->chain_ok = false;
  int64_t* d_map = nullptr;
  float* d_tmp = nullptr;
  CU_TRY(dev_alloc(&d_map, static_cast<size_t>(rows) * sizeof(int64_t), h->stream));
  cudaError_t e = cudaMemcpyAsync(d_map, row_map, static_cast<size_t>(rows) * sizeof(int64_t),
                                  cudaMemcpyHostToDevice, h->stream);
  const int64_t lines = h->c_row_major ? out_rows : n;
  float* dst = C;
  int64_t dst_ld = ld;
  if (e == cudaSuccess && !on_device) {
    // host destination: scatter into a compact device image of the whole output first (rows no
    // block-row of this handle maps to come back as zeros)
    dst_ld = width;
    const size_t bytes = static_cast<size_t>(lines) * width * sizeof(float);
    e = dev_alloc(&d_tmp, bytes, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_tmp, 0, bytes, h->stream);
    dst = d_tmp;
  }
  if (e == cudaSuccess)
    e = permute_rows(h->d_C, h->c_row_major ? h->ldc : 1, h->c_row_major ? 1 : h->ldc, dst,
                     h->c_row_major ? dst_ld : 1, h->c_row_major ? 1 : dst_ld, d_map, rows, n, h->stream);
  if (e == cudaSuccess && !on_device)
    e = cudaMemcpy2DAsync(C, ld * sizeof(float), d_tmp, width * sizeof(float), width * sizeof(float), lines,
                          cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  dev_free(d_map, h->stream);
  dev_free(d_tmp, h->stream);
  if (e != cudaSuccess) return fail_cuda(e, "permuted read-back of C");
  return SPARTA_OK;
}

static int launch(sparta_handle* h, unsigned long long* trace = nullptr, int trace_worker = 0,
                  int trace_cap = 0) {
  if (!h) return fail(SPARTA_ERR_INVALID, "NULL handle");
  if (h->n == 0) return fail(SPARTA_ERR_STATE, "set_B must be called before run");
  CU_TRY(cudaSetDevice(h->device));
  if (h->kind == 1) {
    if (trace) return fail(SPARTA_ERR_INVALID, "CSR handles have no worker timeline");
    CsrParams c;
    memset(&c, 0, sizeof(c));
    c.rowptr = h->d_rowptr; c.colind = h->d_colind; c.val = h->d_val; c.row_order = h->d_row_order;
    c.B = h->d_B; c.C = h->d_C;
    c.c_sr = h->c_row_major ? h->ldc : 1;
    c.c_sj = h->c_row_major ? 1 : h->ldc;
    c.rows = h->rows;
    c.heavy_rows = h->heavy_rows;
    c.n = static_cast<int32_t>(h->n);
    c.ldn = static_cast<int32_t>(h->ldk);
    c.accumulate = h->accumulate;
    const cudaError_t e = spmm_csr_launch(c, h->sopt.precision, h->stream);
    if (e != cudaSuccess) return fail_cuda(e, "CSR kernel launch");
    if (h->rows > 0) ++h->launches;
    return SPARTA_OK;
  }
  if (h->gather_work > 0) {
    // the short block-rows of a hybrid handle (disjoint rows of C, same stream)
    if (trace) return fail(SPARTA_ERR_INVALID, "worker timelines need a handle without gather rows (gather_max_height = -1)");
    for (size_t pi = 0; pi < h->gather_passes.size(); ++pi) {
      const sparta_handle::GatherPassDev& d = h->gather_passes[pi];
      if (d.work == 0) continue;
      CsrParams c;
      memset(&c, 0, sizeof(c));
      c.rowptr = d.beg; c.rowend = d.end; c.colind = h->d_colind; c.val = h->d_val; c.row_order = d.order;
      c.B = h->d_B2; c.C = h->d_C;
      c.c_sr = h->c_row_major ? h->ldc : 1;
      c.c_sj = h->c_row_major ? 1 : h->ldc;
      c.rows = d.work;
      c.heavy_rows = d.heavy;
      c.n = static_cast<int32_t>(h->n);
      c.ldn = static_cast<int32_t>(h->ldn2);
      c.accumulate = pi == 0 ? h->accumulate : 1;   // pass 0 writes the rows, the later ranges of k add to them
      const cudaError_t e = spmm_csr_launch(c, h->sopt.precision, h->stream);
      if (e != cudaSuccess) return fail_cuda(e, "gather kernel launch");
      ++h->launches;
    }
  }
  if (h->as.grid == 0) return SPARTA_OK;  // empty shard / nothing for the tensor cores
  SpmmParams p;
  memset(&p, 0, sizeof(p));
  p.items = h->d_items; p.cta_ptr = h->d_cta_ptr; p.cta_items = h->d_cta_items;
  p.srows = h->d_srows; p.segs = h->d_segs; p.chunks = h->d_chunks; p.a_packed = h->d_a;
  p.tables = reinterpret_cast<const uint8_t*>(h->d_tables);
  p.C = h->d_C;
  p.c_sr = h->c_row_major ? h->ldc : 1;
  p.c_sj = h->c_row_major ? 1 : h->ldc;
  p.n = static_cast<int32_t>(h->n);
  p.accumulate = h->accumulate;
  p.pair = h->st.pair;
  p.producers = h->producers;
  p.a_slot_bytes = h->a_slot_bytes;
  p.trace = trace;
  p.trace_worker = trace_worker;
  p.trace_cap = trace_cap;
  p.kind_tf32 = h->sopt.precision == PREC_TF32;
  p.panel_stages = h->panel_stages;
  p.a_ring_bytes = h->a_ring_bytes;
  // bounded chains: 256 working columns + 256 master columns, one accumulator stage
  p.master_col = h->st.master_col;
  p.tiles = h->st.tiles;
  p.acc_stages = (h->st.master_col > 0 || h->st.tiles > 1) ? 1 : 512 / h->st.acc_cols;
  p.acc_stage_cols = h->st.acc_cols;
  const char* err = "";
  if (!h->as.zero_jobs.empty() && !h->accumulate) {
    // split pieces add partial sums: their C tiles start from zero (C := A*B semantics); the
    // kernel zeroes them itself and every CTA of the grid reports in on the counter.  The counter
    // target is a per-launch kernel parameter, so a captured launch must not be replayed.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CU_TRY(cudaStreamIsCapturing(h->stream, &cap));
    if (cap != cudaStreamCaptureStatusNone)
      return fail(SPARTA_ERR_STATE, "a handle whose plan has split pieces cannot be captured into a CUDA graph; "
                                    "create it with split_k = 1");
    if (!h->d_sync) {
      CU_TRY(dev_alloc(&h->d_sync, sizeof(unsigned long long), h->stream));
      CU_TRY(cudaMemsetAsync(h->d_sync, 0, sizeof(unsigned long long), h->stream));
      h->sync_total = 0;
    }
    h->sync_total += static_cast<unsigned long long>(h->as.grid);
    p.zero_jobs = h->d_zero_jobs;
    p.n_zero_jobs = static_cast<int32_t>(h->as.zero_jobs.size());
    p.sync_counter = h->d_sync;
    p.sync_target = h->sync_total;
  }
  // Back-to-back multiplies of one handle are chained with programmatic dependent launch.  Measured
  // (profiles/r2_pdl.md): letting a whole item's MMAs overlap the previous grid's tail (mode 1) helps
  // mid-size shards (a quarter of config #3: 0.860 -> 0.790 ms) but costs config #3 itself 15 % -- the
  // workers of a team start out of step and re-fetch their A images; overlapping only the launch, the
  // prologue and the first stages' copies (mode 2, the default) is 2-4 % faster everywhere tried.
  // SPARTA_PDL_MODE = 0 / 1 / 2 overrides.
  static const int chain_mode = getenv("SPARTA_PDL_MODE") ? atoi(getenv("SPARTA_PDL_MODE")) : 2;
  const bool chain = h->chain_ok && chain_mode > 0 && trace == nullptr && h->gather_work == 0;
  p.chain_wait_mma = chain_mode != 1;
  cudaError_t e = spmm_launch(p, h->d_B, h->cols, h->ldk, h->sopt.precision, h->as.grid, h->stream, &err, chain);
  if (e != cudaSuccess) {
    if (p.n_zero_jobs) h->sync_total -= static_cast<unsigned long long>(h->as.grid);   // nothing ran
    h->chain_ok = false;
    return fail_cuda(e, err);
  }
  h->chain_ok = trace == nullptr;
  ++h->launches;
  return SPARTA_OK;
}

int sparta_run_async(sparta_handle* h) { return launch(h); }

int sparta_run_traced(sparta_handle* h, int32_t worker, uint64_t* records, int64_t capacity) {
  if (!h || !records || capacity <= 0 || capacity > (1 << 24))
    return fail(SPARTA_ERR_INVALID, "NULL handle/records or invalid capacity");
  if (h->n == 0) return fail(SPARTA_ERR_STATE, "set_B must be called before run");
  CU_TRY(cudaSetDevice(h->device));
  const size_t bytes = static_cast<size_t>(capacity) * 4 * 2 * 2 * sizeof(uint64_t);
  unsigned long long* d = nullptr;
  CU_TRY(dev_alloc(&d, bytes, h->stream));
  cudaError_t e = cudaMemsetAsync(d, 0, bytes, h->stream);
  int rc = SPARTA_OK;
  if (e == cudaSuccess) rc = launch(h, d, worker, static_cast<int>(capacity));
  if (e == cudaSuccess && rc == SPARTA_OK)
    e = cudaMemcpyAsync(records, d, bytes, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && rc == SPARTA_OK) e = cudaStreamSynchronize(h->stream);
  dev_free(d, h->stream);
  if (rc) return rc;
  if (e != cudaSuccess) return fail_cuda(e, "traced run");
  return SPARTA_OK;
}

int sparta_synchronize(sparta_handle* h) {
  if (!h) return fail(SPARTA_ERR_INVALID, "NULL handle");
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  release_staged(h);
  return SPARTA_OK;
}

int sparta_run(sparta_handle* h, float* dt_ms) {
  if (!h) return fail(SPARTA_ERR_INVALID, "NULL handle");
  if (h->n == 0) return fail(SPARTA_ERR_STATE, "set_B must be called before run");
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaEventRecord(h->ev0, h->stream));
  const int rc = launch(h);
  if (rc) return rc;
  CU_TRY(cudaEventRecord(h->ev1, h->stream));
  CU_TRY(cudaEventSynchronize(h->ev1));
  CU_TRY(cudaGetLastError());
  float ms = 0;
  CU_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (dt_ms) *dt_ms = ms;
  return SPARTA_OK;
}

void* sparta_C_device_ptr(sparta_handle* h) { return h ? h->d_C : nullptr; }
int64_t sparta_C_device_ld(sparta_handle* h) { return h ? h->ldc : 0; }
void* sparta_stream(sparta_handle* h) { return h ? h->stream : nullptr; }

static void fill_stats(const Structure& st, const Assignment& as, int64_t cols, int64_t block_rows,
                       int panel_stages, int a_ring_bytes, sparta_stats* s) {
  memset(s, 0, sizeof(*s));
  s->rows = st.rows; s->cols = cols; s->block_rows = block_rows;
  s->nz_blocks = st.n_blocks; s->nztot = st.nztot;
  s->segments = static_cast<int64_t>(st.segs.size());
  s->super_rows = static_cast<int64_t>(st.srows.size());
  s->chunks = static_cast<int64_t>(st.chunks.size());
  s->items = static_cast<int64_t>(as.items.size());
  s->a_packed_bytes = static_cast<int64_t>(st.a_bytes);
  s->grid = as.grid;
  s->smem_bytes = spmm_smem_bytes(panel_stages, a_ring_bytes, st.tiles);
  s->sched_imbalance = as.mean_cta_cost > 0 ? as.max_cta_cost / as.mean_cta_cost : 1.0;
  s->team = as.team;
  s->cta_pair = st.pair;
  s->wide_tiles = st.tiles;
  s->split_pieces = as.split_pieces;
  s->zero_tiles = static_cast<int32_t>(as.zero_jobs.size());
  s->sched_max_cycles = as.max_cta_cost;
}

int sparta_get_stats(sparta_handle* h, sparta_stats* out) {
  if (!h || !out) return fail(SPARTA_ERR_INVALID, "NULL handle or stats");
  fill_stats(h->st, h->as, h->cols, h->block_rows, h->panel_stages, h->a_ring_bytes, out);
  if (h->kind == 1) {
    out->rows = h->rows;
    out->nztot = h->nnz;
    out->smem_bytes = 0;
    out->grid = static_cast<int32_t>(std::min<int64_t>((h->rows + 7) / 8, INT32_MAX));
    out->b_bytes = h->cols * h->ldk * (h->sopt.precision == PREC_TF32 ? 4 : 2);
  } else {
    out->b_bytes = static_cast<int64_t>(h->n) * h->ldk * prec_esize(h->sopt.precision);
  }
  out->c_bytes = h->c_row_major ? h->rows * h->n * 4 : h->ldc * h->n * 4;
  float up = 0;
  if (cudaSetDevice(h->device) == cudaSuccess && cudaEventSynchronize(h->up1) == cudaSuccess &&
      cudaEventElapsedTime(&up, h->up0, h->up1) == cudaSuccess)
    out->upload_ms = up;
  else
    cudaGetLastError();
  out->kernel_launches = h->launches;
  if (h->kind == 0) {
    out->gather_rows = h->gather_rows;
    out->gather_nnz = h->gather_work > 0 ? h->nnz : 0;
    out->block_rows = h->block_rows;
    if (h->gather_work > 0) out->b_bytes += h->cols * h->ldn2 * prec_esize(h->sopt.precision);
  }
  return SPARTA_OK;
}

int sparta_destroy(sparta_handle* h) {
  free_handle(h);
  return SPARTA_OK;
}

// One-shot data flow shared by the three formats: nothing synchronises between the upload of A,
// the upload of B, the kernel and the download of C -- they are one in-order sequence on the
// handle's stream, and the host scheduler overlaps the A upload (create_common).
}  // extern "C"
template <class CreateFn>
static int one_shot(const char* name, CreateFn create, const float* B, int64_t ldb, int64_t n, float* C,
                    int64_t ldc, float* dt_ms) {
  sparta_handle* h = nullptr;
  const bool timing = getenv("SPARTA_TIMING") != nullptr;   // phase breakdown of the call on stderr
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto t0 = now();
  int rc = guarded([&] { return create(&h); });
  if (rc) return rc;
  const auto t1 = now();
  rc = set_b_impl(h, B, ldb, n, 0, true);
  const auto t2 = now();
  if (!rc) {
    cudaError_t e = cudaEventRecord(h->ev0, h->stream);
    if (e == cudaSuccess) { rc = sparta_run_async(h); e = cudaEventRecord(h->ev1, h->stream); }
    if (e != cudaSuccess) rc = fail_cuda(e, "event record");
  }
  const auto t3 = now();
  if (!rc) rc = sparta_get_C(h, C, ldc, 0);   // synchronises the stream
  const auto t4 = now();
  if (!rc && dt_ms) {
    cudaError_t e = cudaEventElapsedTime(dt_ms, h->ev0, h->ev1);
    if (e != cudaSuccess) rc = fail_cuda(e, "cudaEventElapsedTime");
  }
  if (rc) cudaStreamSynchronize(h->stream);   // host buffers must be idle before we return
  const std::string keep = g_last_error;
  sparta_destroy(h);
  if (timing)
    fprintf(stderr, "%s: enqueue create %.1f ms (host schedule overlapped with the A upload), set_B %.1f, run %.1f, "
            "wait + get_C %.1f, destroy %.1f\n", name, ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, now()));
  if (rc) g_last_error = keep;
  return rc;
}
extern "C" {

int sparta_vbr_spmm(int64_t rows, int64_t cols, int64_t block_rows, int64_t block_col_size,
                    const int64_t* row_part, const int64_t* nzcount, const int64_t* jab,
                    const float* mab, const float* B, int64_t ldb, int64_t n, float* C,
                    int64_t ldc, int precision, float* dt_ms) {
  sparta_options o;
  memset(&o, 0, sizeof(o));
  o.struct_size = sizeof(o);
  o.precision = precision;
  o.n_hint = static_cast<int32_t>(std::min<int64_t>(n, INT32_MAX));
  return one_shot("sparta_vbr_spmm", [&](sparta_handle** h) {
    return vbr_create_impl(h, rows, cols, block_rows, block_col_size, row_part, nzcount, jab, mab, &o, true);
  }, B, ldb, n, C, ldc, dt_ms);
}

int sparta_csr_vbr_spmm(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind, const float* val,
                        const int64_t* grouping, int64_t block_col_size, int64_t row_block_size,
                        int32_t force_fixed_size, const float* B, int64_t ldb, int64_t n, float* C, int64_t ldc,
                        int precision, float* dt_ms) {
  sparta_options o;
  memset(&o, 0, sizeof(o));
  o.struct_size = sizeof(o);
  o.precision = precision;
  o.n_hint = static_cast<int32_t>(std::min<int64_t>(n, INT32_MAX));
  // B starts crossing PCIe BEFORE the host builds the index arrays and the tile schedule (tens of ms on
  // the CPU): a staging buffer on a side stream of the current device, handed to set_B as a device
  // operand once the handle exists.
  const bool timing = getenv("SPARTA_TIMING") != nullptr;
  const auto tt0 = std::chrono::steady_clock::now();
  auto ms_since = [&](std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
  };
  float* d_stage = nullptr;
  cudaStream_t side = nullptr;
  cudaEvent_t landed = nullptr;
  const int64_t b_cols = force_fixed_size ? ((cols - 1) / block_col_size + 1) * block_col_size : cols;
  if (!B || n <= 0 || ldb < b_cols) return fail(SPARTA_ERR_INVALID, "invalid B");
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return fail(SPARTA_ERR_NO_DEVICE, "no CUDA device visible (libsparta_b200 has no CPU path)"); }
  cudaError_t e = pool_prepare(dev);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&landed, cudaEventDisableTiming);
  if (e == cudaSuccess) e = dev_alloc(&d_stage, static_cast<size_t>(n) * b_cols * sizeof(float), side);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(d_stage, b_cols * sizeof(float), B, ldb * sizeof(float), b_cols * sizeof(float), n,
                          cudaMemcpyHostToDevice, side);
  if (e == cudaSuccess) e = cudaEventRecord(landed, side);
  int rc = SPARTA_OK;
  if (e != cudaSuccess) rc = fail_cuda(e, "staging of B");
  sparta_handle* h = nullptr;
  if (!rc)
    rc = guarded([&] {
      return vbr_create_from_csr_impl(&h, rows, cols, rowptr, colind, val, grouping, block_col_size, row_block_size,
                                      force_fixed_size, &o, true, nullptr);
    });
  if (!rc) {
    e = cudaStreamWaitEvent(h->stream, landed, 0);
    if (e != cudaSuccess) rc = fail_cuda(e, "cudaStreamWaitEvent");
  }
  const double t_create = ms_since(tt0);
  if (!rc) rc = set_b_impl(h, d_stage, b_cols, n, 1, true);
  if (!rc) {
    e = cudaEventRecord(h->ev0, h->stream);
    if (e == cudaSuccess) { rc = sparta_run_async(h); e = cudaEventRecord(h->ev1, h->stream); }
    if (e != cudaSuccess && !rc) rc = fail_cuda(e, "event record");
  }
  const double t_enq = ms_since(tt0);
  if (!rc) rc = sparta_get_C(h, C, ldc, 0);   // synchronises the stream
  const double t_done = ms_since(tt0);
  if (!rc && dt_ms) {
    e = cudaEventElapsedTime(dt_ms, h->ev0, h->ev1);
    if (e != cudaSuccess) rc = fail_cuda(e, "cudaEventElapsedTime");
  }
  const std::string keep = g_last_error;
  if (h) { cudaStreamSynchronize(h->stream); sparta_destroy(h); }
  if (side) {
    cudaStreamSynchronize(side);
    if (d_stage) cudaFreeAsync(d_stage, side);
    cudaStreamSynchronize(side);
    cudaStreamDestroy(side);
  }
  if (landed) cudaEventDestroy(landed);
  if (timing)
    fprintf(stderr, "sparta_csr_vbr_spmm: B staged + handle from CSR %.1f ms, set_B + run enqueued %.1f, wait + C down %.1f, "
            "teardown %.1f\n", t_create, t_enq - t_create, t_done - t_enq, ms_since(tt0) - t_done);
  if (rc) g_last_error = keep;
  return rc;
}

int sparta_vbr_spmm_BA(int64_t rows, int64_t cols, int64_t block_rows, int64_t block_col_size,
                       const int64_t* row_part, const int64_t* nzcount, const int64_t* jab,
                       const float* mab, const float* B, int64_t ldb, int64_t n, float* C,
                       int64_t ldc, int precision, float* dt_ms) {
  sparta_options o;
  memset(&o, 0, sizeof(o));
  o.struct_size = sizeof(o);
  o.precision = precision;
  o.n_hint = static_cast<int32_t>(std::min<int64_t>(n, INT32_MAX));
  return one_shot("sparta_vbr_spmm_BA", [&](sparta_handle** h) {
    return vbr_create_ba_impl(h, rows, cols, block_rows, block_col_size, row_part, nzcount, jab, mab, &o, true);
  }, B, ldb, n, C, ldc, dt_ms);
}

int sparta_bellpack_spmm(int64_t rows, int64_t cols, int64_t ell_blocksize,
                         int64_t ellColInd_rows, int64_t ellColInd_cols,
                         const int64_t* ellColInd, const float* ellValues, const float* B,
                         int64_t ldb, int64_t n, float* C, int64_t ldc, int precision,
                         float* dt_ms) {
  sparta_options o;
  memset(&o, 0, sizeof(o));
  o.struct_size = sizeof(o);
  o.precision = precision;
  o.n_hint = static_cast<int32_t>(std::min<int64_t>(n, INT32_MAX));
  return one_shot("sparta_bellpack_spmm", [&](sparta_handle** h) {
    return bellpack_create_impl(h, rows, cols, ell_blocksize, ellColInd_rows, ellColInd_cols, ellColInd,
                                ellValues, &o, true);
  }, B, ldb, n, C, ldc, dt_ms);
}

int sparta_csr_spmm(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                    const float* val, const float* B, int64_t ldb, int64_t n, float* C, int64_t ldc,
                    int precision, float* dt_ms) {
  sparta_options o;
  memset(&o, 0, sizeof(o));
  o.struct_size = sizeof(o);
  o.precision = precision;
  return one_shot("sparta_csr_spmm", [&](sparta_handle** h) {
    return csr_create_impl(h, rows, cols, rowptr, colind, val, &o, true);
  }, B, ldb, n, C, ldc, dt_ms);
}

int sparta_release_workspace(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return SPARTA_OK; }
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < ndev && d < 64; ++d) {
    bool ready;
    { std::lock_guard<std::mutex> lock(g_pool_mutex); ready = g_pool_ready[d]; }
    if (!ready) continue;
    cudaMemPool_t pool;
    if (cudaSetDevice(d) != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, d) != cudaSuccess) continue;
    cudaDeviceSynchronize();
    cudaMemPoolTrimTo(pool, 0);
  }
  cudaSetDevice(cur);
  cudaGetLastError();
  pinned_trim();
  return SPARTA_OK;
}

}  // extern "C"
// Modelled time, in SM cycles, of the shard [lo, hi): the slowest worker of its tile schedule plus the
// gather rows that run before it on the same stream.
static const char* shard_model_cycles(int64_t block_rows, int64_t block_col_size, const int64_t* row_part,
                                      const int64_t* nzcount, const int64_t* jab, int64_t lo, int64_t hi, int64_t cols,
                                      int64_t n, const sparta_options& o, const ScheduleOptions& so, double* out) {
  *out = 0;
  if (hi <= lo) return "";
  BlockRows br;
  int64_t src_lo = 0, src_hi = 0;
  const char* e = blockrows_from_vbr(block_rows, block_col_size, row_part, nzcount, jab, lo, hi, &br, &src_lo, &src_hi);
  Structure st;
  Assignment as;
  BlockRows fused, tall;
  double gather_nnz = 0;
  const int gather_h = o.gather_max_height < 0 ? 0 : (o.gather_max_height == 0 ? 7 : o.gather_max_height);
  const BlockRows* view = &br;
  if (!*e && split_short_view(br, gather_h, &tall, &gather_nnz)) view = &tall;
  const bool use_fused = !*e && o.fuse_rows != 1 && fuse_short_block_rows(*view, 16, &fused);
  ScheduleOptions so2 = so;
  sparta_options o2 = o;
  o2.n_hint = static_cast<int32_t>(std::min<int64_t>(n, INT32_MAX));
  if (!*e) e = build_structure_choosing_tiles(use_fused ? fused : *view, o2, &so2, &st);
  if (!*e) e = build_assignment(st, so2, n, cols, &as);
  if (*e) return e;
  // the gather rows: one row of B (n elements) per nonzero through the L2 at ~40 bytes per clock and SM
  // (HBM speed when a column tile's slab of B, cols x 256 elements, is larger than the L2)
  const double esb = so.precision == PREC_TF32 ? 4.0 : 2.0;
  const double slab = static_cast<double>(cols) * 256.0 * esb;
  const double rate = so.num_ctas * (slab > 100e6 ? 20.0 : 40.0);
  *out = as.max_cta_cost + gather_nnz * static_cast<double>(n) * esb / rate;
  return "";
}

static ScheduleOptions schedule_options_of(const sparta_options& o) {
  ScheduleOptions so;
  so.precision = o.precision;
  so.seg_rows = o.seg_rows;
  so.acc_cols = o.acc_cols;
  so.num_ctas = o.num_ctas > 0 ? o.num_ctas : default_grid_ctas();
  so.pair = o.cta_pair != 1;
  so.sort_rows = o.row_order != 1;
  so.l2_slab_bytes = static_cast<int64_t>(o.l2_slab_mb) << 20;
  so.max_chain = o.max_chain;
  so.split = o.split_k;
  return so;
}

extern "C" int sparta_partition_model_times(int64_t rows, int64_t cols, int64_t block_rows, int64_t block_col_size,
                                            const int64_t* row_part, const int64_t* nzcount, const int64_t* jab, int64_t n,
                                            const sparta_options* opt, int32_t parts, const int64_t* cuts, double* cycles) {
  if (block_rows < 0 || parts <= 0 || !cuts || !cycles || cols <= 0 || block_col_size <= 0 || n <= 0 ||
      (block_rows && (!row_part || !nzcount)))
    return fail(SPARTA_ERR_INVALID, "invalid partition request");
  (void)rows;
  sparta_options o;
  resolve_options(opt, &o);
  const ScheduleOptions so = schedule_options_of(o);
  for (int i = 0; i < parts; ++i) {
    if (cuts[i] < 0 || cuts[i + 1] > block_rows || cuts[i] > cuts[i + 1]) return fail(SPARTA_ERR_INVALID, "cuts out of range");
    const char* e = shard_model_cycles(block_rows, block_col_size, row_part, nzcount, jab, cuts[i], cuts[i + 1], cols, n, o, so,
                                       &cycles[i]);
    if (*e) return fail(SPARTA_ERR_INVALID, e);
  }
  return SPARTA_OK;
}

extern "C" {

// Contiguous block-row ranges balanced on the MODELLED kernel time of each shard instead of its
// nonzero-block area: sparse block-rows cost more per FLOP (a B panel is fetched per chunk however
// few rows share it) and every work item pays a fixed drain, so equal areas are not equal times --
// on the bench matrix the sparse tail shard of an area partition runs 1.3-1.4x longer than the
// dense head.  Fixed point: start from the area partition; build every shard's schedule; spread
// its modelled time (the slowest worker's cycles) over its block-rows in proportion to their
// current weight; re-cut on the new weights; keep the best of a few rounds.
static int partition_modelled_impl(int64_t rows, int64_t cols, int64_t block_rows,
                                   int64_t block_col_size, const int64_t* row_part,
                                   const int64_t* nzcount, const int64_t* jab, int64_t n,
                                   const sparta_options* opt, int32_t parts, const double* time_scale,
                                   int64_t* cuts);

int sparta_partition_block_rows_modelled(int64_t rows, int64_t cols, int64_t block_rows,
                                         int64_t block_col_size, const int64_t* row_part,
                                         const int64_t* nzcount, const int64_t* jab, int64_t n,
                                         const sparta_options* opt, int32_t parts, int64_t* cuts) {
  return guarded([&] {
    return partition_modelled_impl(rows, cols, block_rows, block_col_size, row_part, nzcount, jab, n, opt, parts, nullptr, cuts);
  });
}

// The same with MEASURED feedback: time_scale[b] = (measured / modelled kernel time) of the shard
// block-row b belonged to in an earlier partition.  Every shard's modelled time is multiplied by
// the weight-averaged scale of its block-rows before the cuts are balanced, which removes what
// the cost model gets systematically wrong about a region of the matrix.
int sparta_partition_block_rows_measured(int64_t rows, int64_t cols, int64_t block_rows,
                                         int64_t block_col_size, const int64_t* row_part,
                                         const int64_t* nzcount, const int64_t* jab, int64_t n,
                                         const sparta_options* opt, int32_t parts,
                                         const double* time_scale, int64_t* cuts) {
  if (!time_scale) return fail(SPARTA_ERR_INVALID, "time_scale is NULL");
  return guarded([&] {
    return partition_modelled_impl(rows, cols, block_rows, block_col_size, row_part, nzcount, jab, n, opt, parts, time_scale, cuts);
  });
}

static int partition_modelled_impl(int64_t rows, int64_t cols, int64_t block_rows,
                                   int64_t block_col_size, const int64_t* row_part,
                                   const int64_t* nzcount, const int64_t* jab, int64_t n,
                                   const sparta_options* opt, int32_t parts, const double* time_scale,
                                   int64_t* cuts) {
  if (block_rows < 0 || parts <= 0 || !cuts || cols <= 0 || block_col_size <= 0 || n <= 0 ||
      (block_rows && (!row_part || !nzcount)))
    return fail(SPARTA_ERR_INVALID, "invalid partition request");
  (void)rows;
  sparta_options o;
  resolve_options(opt, &o);
  ScheduleOptions so;
  so.precision = o.precision;
  so.seg_rows = o.seg_rows;
  so.acc_cols = o.acc_cols;
  so.num_ctas = o.num_ctas > 0 ? o.num_ctas : default_grid_ctas();
  so.pair = o.cta_pair != 1;
  so.sort_rows = o.row_order != 1;
  so.l2_slab_bytes = static_cast<int64_t>(o.l2_slab_mb) << 20;
  so.max_chain = o.max_chain;
  so.split = o.split_k;
  std::vector<double> weight(block_rows);
  for (int64_t b = 0; b < block_rows; ++b)
    weight[b] = 1.0 + static_cast<double>(nzcount[b]) * (row_part[b + 1] - row_part[b]);
  std::vector<int64_t> cur(parts + 1), best(parts + 1);
  double best_worst = -1;
  for (int round = 0; round < 6; ++round) {
    // cut on the prefix sums of the weights
    std::vector<double> prefix(block_rows + 1, 0.0);
    for (int64_t b = 0; b < block_rows; ++b) prefix[b + 1] = prefix[b] + weight[b];
    cur[0] = 0;
    for (int i = 1; i < parts; ++i) {
      const double target = prefix[block_rows] * i / parts;
      int64_t b = std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin();
      if (b > 0 && target - prefix[b - 1] < prefix[b] - target) --b;
      cur[i] = std::min(std::max(b, cur[i - 1]), block_rows);
    }
    cur[parts] = block_rows;
    // modelled time of every shard
    std::vector<double> t(parts, 0.0);
    double worst = 0;
    for (int i = 0; i < parts; ++i) {
      if (cur[i + 1] <= cur[i]) continue;
      const char* e = shard_model_cycles(block_rows, block_col_size, row_part, nzcount, jab, cur[i], cur[i + 1], cols, n, o,
                                         so, &t[i]);
      if (*e) return fail(SPARTA_ERR_INVALID, e);
      if (time_scale) {
        double ws = 0, w = 0;
        for (int64_t b = cur[i]; b < cur[i + 1]; ++b) { ws += weight[b] * time_scale[b]; w += weight[b]; }
        if (w > 0) t[i] *= ws / w;
      }
      worst = std::max(worst, t[i]);
    }
    if (best_worst < 0 || worst < best_worst) { best_worst = worst; best = cur; }
    for (int i = 0; i < parts; ++i) {
      double wsum = 0;
      for (int64_t b = cur[i]; b < cur[i + 1]; ++b) wsum += weight[b];
      if (wsum <= 0 || t[i] <= 0) continue;
      const double scale = t[i] / wsum;
      for (int64_t b = cur[i]; b < cur[i + 1]; ++b) weight[b] *= scale;
    }
  }
  for (int i = 0; i <= parts; ++i) cuts[i] = best[i];
  return SPARTA_OK;
}

int sparta_partition_block_rows(int64_t block_rows, const int64_t* row_part,
                                const int64_t* nzcount, int32_t parts, int64_t* cuts) {
  if (block_rows < 0 || parts <= 0 || !cuts || (block_rows && (!row_part || !nzcount)))
    return fail(SPARTA_ERR_INVALID, "invalid partition request");
  partition_block_rows(block_rows, row_part, nzcount, parts, cuts);
  return SPARTA_OK;
}

int sparta_host_blocking(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                         int32_t algo, float tau, int64_t block_col_size, int64_t row_block_size,
                         int32_t sim_measure, int32_t use_pattern, int32_t use_groups,
                         int32_t force_fixed_size, int32_t flags, int64_t* grouping,
                         sparta_blocking_stats* stats) {
  if (rows < 0 || !rowptr || (rows && !grouping) || (rows && rowptr[rows] && !colind))
    return fail(SPARTA_ERR_INVALID, "NULL CSR or grouping array");
  BlockingParams p;
  p.algo = algo; p.tau = tau; p.block_col_size = block_col_size; p.row_block_size = row_block_size;
  p.sim_measure = sim_measure; p.use_pattern = use_pattern != 0; p.use_groups = use_groups != 0;
  p.force_fixed_size = force_fixed_size != 0; p.force_list_model = (flags & 1) != 0;
  BlockingStats st;
  const auto t0 = std::chrono::steady_clock::now();
  const char* e = host_blocking(rows, cols, rowptr, colind, p, grouping, &st);
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  if (stats) {
    stats->comparison_counter = st.comparisons;
    stats->merge_counter = st.merges;
    stats->average_merge_tau = st.average_merge_tau;
    stats->average_row_distance = st.average_row_distance;
    stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return SPARTA_OK;
}

int sparta_grouping_save(const char* path, int64_t rows, const int64_t* grouping, uint64_t key) {
  if (!path || rows < 0 || (rows && !grouping)) return fail(SPARTA_ERR_INVALID, "NULL path or grouping");
  const char* e = grouping_save(path, rows, grouping, key, "");
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  return SPARTA_OK;
}

int sparta_grouping_load(const char* path, int64_t rows, int64_t* grouping, uint64_t key) {
  if (!path || rows < 0 || (rows && !grouping)) return fail(SPARTA_ERR_INVALID, "NULL path or grouping");
  const char* e = grouping_load(path, rows, grouping, key);
  if (*e) return fail(SPARTA_ERR_INVALID, std::string(e) == "miss" ? "no grouping file with this key" : e);
  return SPARTA_OK;
}

uint64_t sparta_blocking_key(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                             int32_t algo, float tau, int64_t block_col_size, int64_t row_block_size,
                             int32_t sim_measure, int32_t use_pattern, int32_t use_groups,
                             int32_t force_fixed_size) {
  if (rows < 0 || !rowptr || (rows && rowptr[rows] && !colind)) return 0;
  return blocking_key(rows, cols, rowptr, colind, algo, tau, block_col_size, row_block_size, sim_measure,
                      use_pattern, use_groups, force_fixed_size);
}

int sparta_host_blocking_cached(const char* cache_dir, int64_t rows, int64_t cols, const int64_t* rowptr,
                                const int64_t* colind, int32_t algo, float tau, int64_t block_col_size,
                                int64_t row_block_size, int32_t sim_measure, int32_t use_pattern,
                                int32_t use_groups, int32_t force_fixed_size, int32_t flags,
                                int64_t* grouping, sparta_blocking_stats* stats, int32_t* hit) {
  if (hit) *hit = 0;
  if (!cache_dir || !*cache_dir)
    return sparta_host_blocking(rows, cols, rowptr, colind, algo, tau, block_col_size, row_block_size, sim_measure,
                                use_pattern, use_groups, force_fixed_size, flags, grouping, stats);
  if (rows < 0 || !rowptr || (rows && !grouping) || (rows && rowptr[rows] && !colind))
    return fail(SPARTA_ERR_INVALID, "NULL CSR or grouping array");
  const auto t0 = std::chrono::steady_clock::now();
  const uint64_t key = blocking_key(rows, cols, rowptr, colind, algo, tau, block_col_size, row_block_size,
                                    sim_measure, use_pattern, use_groups, force_fixed_size);
  char name[96];
  snprintf(name, sizeof(name), "/grouping_%016llx.g", static_cast<unsigned long long>(key));
  const std::string path = std::string(cache_dir) + name;
  const char* e = grouping_load(path.c_str(), rows, grouping, key);
  if (!*e) {
    if (hit) *hit = 1;
    if (stats) {   // the merge statistics are not stored: a hit reports the lookup time only
      memset(stats, 0, sizeof(*stats));
      stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return SPARTA_OK;
  }
  if (std::string(e) != "miss") return fail(SPARTA_ERR_INVALID, e);
  const int rc = sparta_host_blocking(rows, cols, rowptr, colind, algo, tau, block_col_size, row_block_size,
                                      sim_measure, use_pattern, use_groups, force_fixed_size, flags, grouping, stats);
  if (rc) return rc;
  char note[256];
  snprintf(note, sizeof(note), "rows %lld cols %lld nnz %lld -a %d -t %g -b %lld -B %lld -m %d -p %d -g %d -F %d",
           static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(rowptr[rows]), algo,
           static_cast<double>(tau), static_cast<long long>(block_col_size), static_cast<long long>(row_block_size),
           sim_measure, use_pattern, use_groups, force_fixed_size);
  e = grouping_save(path.c_str(), rows, grouping, key, note);
  if (*e) return fail(SPARTA_ERR_INVALID, e);   // an unwritable cache directory is the caller's mistake
  return SPARTA_OK;
}

int sparta_host_row_order(int64_t rows, const int64_t* rowptr, int32_t mode, uint32_t seed, int64_t* order) {
  if (rows < 0 || !rowptr || (rows && !order)) return fail(SPARTA_ERR_INVALID, "NULL rowptr or order");
  const char* e = host_row_order(rows, rowptr, mode, seed, order);
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  return SPARTA_OK;
}

int sparta_host_permutation(int64_t n, const int64_t* grouping, int64_t* perm) {
  if (n < 0 || (n && (!grouping || !perm))) return fail(SPARTA_ERR_INVALID, "NULL grouping or perm");
  host_permutation(grouping, n, perm);
  return SPARTA_OK;
}

int sparta_host_partition(int64_t n, const int64_t* grouping, int64_t* part, int64_t* part_len) {
  if (n < 0 || !part || !part_len || (n && !grouping)) return fail(SPARTA_ERR_INVALID, "NULL argument");
  *part_len = host_partition(grouping, n, part);
  return SPARTA_OK;
}

int sparta_host_vbr_fill(sparta_host_vbr** out, int64_t rows, int64_t cols, const int64_t* rowptr,
                         const int64_t* colind, const float* val, int32_t pattern_only,
                         const int64_t* grouping, int64_t block_col_size, int64_t row_block_size,
                         int32_t force_fixed_size, int32_t threads) {
  if (!out) return fail(SPARTA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (rows <= 0 || cols <= 0 || !rowptr || !grouping || (rowptr[rows] && !colind) ||
      (!pattern_only && rowptr[rows] && !val))
    return fail(SPARTA_ERR_INVALID, "invalid CSR input");
  sparta_host_vbr* v = new sparta_host_vbr();
  const char* e = host_vbr_fill(rows, cols, rowptr, colind, val, pattern_only != 0, grouping,
                                block_col_size, row_block_size, force_fixed_size != 0,
                                threads > 0 ? threads : 1, &v->v);
  if (*e) { delete v; return fail(SPARTA_ERR_INVALID, e); }
  *out = v;
  return SPARTA_OK;
}

int sparta_host_vbr_get(sparta_host_vbr* v, int64_t* dims, const int64_t** row_part,
                        const int64_t** nzcount, const int64_t** jab, const float** mab) {
  if (!v || !dims) return fail(SPARTA_ERR_INVALID, "NULL argument");
  dims[0] = v->v.rows; dims[1] = v->v.cols; dims[2] = v->v.block_rows; dims[3] = v->v.block_cols;
  dims[4] = v->v.block_col_size; dims[5] = v->v.nztot;
  if (row_part) *row_part = v->v.row_part.data();
  if (nzcount) *nzcount = v->v.nzcount.data();
  if (jab) *jab = v->v.jab.data();
  if (mab) *mab = v->v.mab.data();
  return SPARTA_OK;
}

int sparta_host_vbr_free(sparta_host_vbr* v) { delete v; return SPARTA_OK; }

int sparta_host_bellpack_from_vbr(sparta_host_bell** out, int64_t rows, int64_t cols,
                                  int64_t block_col_size, const int64_t* nzcount,
                                  const int64_t* jab, const float* mab, int32_t threads) {
  if (!out) return fail(SPARTA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!nzcount) return fail(SPARTA_ERR_INVALID, "NULL nzcount");
  sparta_host_bell* b = new sparta_host_bell();
  const char* e = host_bellpack_from_vbr(rows, cols, block_col_size, nzcount, jab, mab,
                                         threads > 0 ? threads : 1, &b->b);
  if (*e) { delete b; return fail(SPARTA_ERR_INVALID, e); }
  *out = b;
  return SPARTA_OK;
}

int sparta_host_bellpack_get(sparta_host_bell* b, int64_t* dims, const int64_t** ellColInd,
                             const float** ellValues) {
  if (!b || !dims) return fail(SPARTA_ERR_INVALID, "NULL argument");
  dims[0] = b->b.blocksize; dims[1] = b->b.ind_rows; dims[2] = b->b.ind_cols;
  if (ellColInd) *ellColInd = b->b.col_ind.data();
  if (ellValues) *ellValues = b->b.values.data();
  return SPARTA_OK;
}

int sparta_host_bellpack_free(sparta_host_bell* b) { delete b; return SPARTA_OK; }

int sparta_vbr_plan_create(sparta_plan** out, int64_t rows, int64_t cols, int64_t block_rows,
                           int64_t block_col_size, const int64_t* row_part,
                           const int64_t* nzcount, const int64_t* jab, int64_t n,
                           const sparta_options* opt) {
  if (!out) return fail(SPARTA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (rows < 0 || cols <= 0 || block_rows < 0 || block_col_size <= 0 || !row_part || (block_rows && !nzcount))
    return fail(SPARTA_ERR_INVALID, "invalid VBR dimensions or NULL index arrays");
  sparta_options o;
  resolve_options(opt, &o);
  BlockRows br;
  int64_t src_lo = 0, src_hi = 0;
  const int64_t lo = o.block_row_begin;
  const int64_t hi = range_end(o, block_rows);
  const char* e = blockrows_from_vbr(block_rows, block_col_size, row_part, nzcount, jab, lo, hi, &br, &src_lo, &src_hi);
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  sparta_plan* p = new sparta_plan();
  p->sopt.precision = o.precision;
  p->sopt.seg_rows = o.seg_rows;
  p->sopt.acc_cols = o.acc_cols;
  p->sopt.num_ctas = o.num_ctas > 0 ? o.num_ctas : default_grid_ctas();
  p->sopt.pair = o.cta_pair != 1;
  p->sopt.sort_rows = o.row_order != 1;
  p->sopt.l2_slab_bytes = static_cast<int64_t>(o.l2_slab_mb) << 20;
  p->sopt.max_chain = o.max_chain;
  p->sopt.split = o.split_k;
  p->cols = cols; p->block_rows = br.count(); p->n = n;
  p->panel_stages = o.panel_stages;
  p->a_ring_bytes = ring_bytes_for(o.panel_stages);
  {
    BlockRows fused;
    const bool use_fused = o.fuse_rows != 1 && fuse_short_block_rows(br, 16, &fused);
    sparta_options o2 = o;
    o2.n_hint = static_cast<int32_t>(std::min<int64_t>(n, INT32_MAX));   // a plan knows its n
    e = build_structure_choosing_tiles(use_fused ? fused : br, o2, &p->sopt, &p->st);
  }
  if (!*e) e = build_assignment(p->st, p->sopt, n, cols, &p->as);
  if (*e) { delete p; return fail(SPARTA_ERR_INVALID, e); }
  *out = p;
  return SPARTA_OK;
}

int sparta_vbr_plan_create_BA(sparta_plan** out, int64_t rows, int64_t cols, int64_t block_rows,
                              int64_t block_col_size, const int64_t* row_part,
                              const int64_t* nzcount, const int64_t* jab, int64_t n,
                              const sparta_options* opt) {
  if (!out) return fail(SPARTA_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (rows <= 0 || cols <= 0 || block_rows < 0 || block_col_size <= 0 || !row_part || (block_rows && !nzcount))
    return fail(SPARTA_ERR_INVALID, "invalid VBR dimensions or NULL index arrays");
  sparta_options o;
  resolve_options(opt, &o);
  const int64_t bc = (cols - 1) / block_col_size + 1;
  BlockRows br;
  int64_t src_hi = 0;
  const char* e = blockrows_from_vbr_transposed(cols, block_rows, block_col_size, row_part, nzcount, jab,
                                                o.block_row_begin, range_end(o, bc),
                                                &br, &src_hi);
  if (*e) return fail(SPARTA_ERR_INVALID, e);
  sparta_plan* p = new sparta_plan();
  p->sopt.precision = o.precision;
  p->sopt.seg_rows = o.seg_rows;
  p->sopt.acc_cols = o.acc_cols;
  p->sopt.num_ctas = o.num_ctas > 0 ? o.num_ctas : default_grid_ctas();
  p->sopt.pair = o.cta_pair != 1;
  p->sopt.sort_rows = o.row_order != 1;
  p->sopt.l2_slab_bytes = static_cast<int64_t>(o.l2_slab_mb) << 20;
  p->sopt.max_chain = o.max_chain;
  p->sopt.split = o.split_k;
  p->cols = rows; p->block_rows = br.count(); p->n = n;
  p->panel_stages = o.panel_stages;
  p->a_ring_bytes = ring_bytes_for(o.panel_stages);
  e = build_structure(br, p->sopt, &p->st);
  if (!*e) e = build_assignment(p->st, p->sopt, n, rows, &p->as);
  if (*e) { delete p; return fail(SPARTA_ERR_INVALID, e); }
  *out = p;
  return SPARTA_OK;
}

int sparta_plan_array(sparta_plan* plan, int32_t which, const void** data, int64_t* count,
                      int32_t* record_bytes) {
  if (!plan || !data || !count || !record_bytes) return fail(SPARTA_ERR_INVALID, "NULL argument");
#define PLAN_ARR(vec, T)                                   \
  do {                                                     \
    *data = (vec).data();                                  \
    *count = static_cast<int64_t>((vec).size());           \
    *record_bytes = static_cast<int32_t>(sizeof(T));       \
    return SPARTA_OK;                                      \
  } while (0)
  switch (which) {
    case 0: PLAN_ARR(plan->st.segs, Segment);
    case 1: PLAN_ARR(plan->st.srows, SuperRow);
    case 2: PLAN_ARR(plan->st.chunks, Chunk);
    case 3: PLAN_ARR(plan->as.items, Item);
    case 4: PLAN_ARR(plan->as.cta_ptr, int32_t);
    case 5: PLAN_ARR(plan->as.cta_items, int32_t);
    case 6: plan->st.merge_jobs(); PLAN_ARR(plan->st.jobs, PackJob);
    case 7: PLAN_ARR(plan->st.tables, uint32_t);
    case 8: PLAN_ARR(plan->as.zero_jobs, ZeroJob);
  }
#undef PLAN_ARR
  return fail(SPARTA_ERR_INVALID, "unknown plan array id");
}

int sparta_plan_stats(sparta_plan* plan, sparta_stats* out) {
  if (!plan || !out) return fail(SPARTA_ERR_INVALID, "NULL argument");
  fill_stats(plan->st, plan->as, plan->cols, plan->block_rows, plan->panel_stages, plan->a_ring_bytes, out);
  return SPARTA_OK;
}

int sparta_plan_destroy(sparta_plan* plan) {
  delete plan;
  return SPARTA_OK;
}

}  // extern "C"
