// Host-side format builders of the product: grouping -> (permutation, partition) ->
// VBR arrays -> Blocked-ELL bundle.  From-scratch, linear-time and multi-threaded,
// but required to reproduce the reference's arrays bit for bit:
//   get_permutation / get_partition   src/general/utilities.cpp:8-43
//   VBR::fill_from_CSR_inplace        src/general/vbr.cpp:135-237  (O(nnz*block_cols) there)
//   prepare_cusparse_BLOCKEDELLPACK   src/cuda/cuda_utilities.cpp:1656-1710
#include <sched.h>
#include "host_formats.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <numeric>
#include <string>
#include <thread>

namespace sparta {

int host_thread_budget(int cap) {
  static const int budget = [] {
    if (const char* e = getenv("SPARTA_THREADS")) {
      const int v = atoi(e);
      if (v > 0) return v;
    }
    long n = 0;
#ifdef __linux__
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
#endif
    if (n <= 0) n = static_cast<long>(std::thread::hardware_concurrency());
    if (n <= 0) n = 8;
    // cgroup v2: "<quota> <period>" or "max <period>"; v1: two files
    double quota = 0;
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
      char q[64];
      long period = 0;
      if (fscanf(f, "%63s %ld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) quota = atof(q) / period;
      fclose(f);
    } else {
      long q = -1, period = 0;
      if (FILE* fq = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { if (fscanf(fq, "%ld", &q) != 1) q = -1; fclose(fq); }
      if (FILE* fp = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(fp, "%ld", &period) != 1) period = 0; fclose(fp); }
      if (q > 0 && period > 0) quota = static_cast<double>(q) / period;
    }
    if (quota >= 1.0 && quota < n) n = static_cast<long>(quota);
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
      const int w = atoi(e);
      if (w > 1) n = std::max(1L, n / w);
    }
    return static_cast<int>(std::max(1L, n));
  }();
  return std::max(1, std::min(budget, cap));
}

// The reference orders rows with std::sort (unstable) keyed on the group id through a
// comparator that takes `int` indices.  The order of rows inside a group is therefore
// whatever libstdc++'s introsort produces; the only way to match it is to run the same
// algorithm with a comparator that answers identically.
void host_permutation(const int64_t* grouping, int64_t n, int64_t* perm) {
  std::iota(perm, perm + n, static_cast<int64_t>(0));
  std::sort(perm, perm + n, [grouping](int a, int b) { return grouping[a] < grouping[b]; });
}

const char* host_row_order(int64_t rows, const int64_t* rowptr, int32_t mode, uint32_t seed, int64_t* order) {
  std::iota(order, order + rows, static_cast<int64_t>(0));
  auto len = [rowptr](int i) { return rowptr[i + 1] - rowptr[i]; };
  if (mode == 1) {
    // the reference's descending comparator is `>=` (csr.cpp:130-133): not a strict weak order, so its
    // std::sort call is undefined behaviour (it segfaults on a 512-row R-MAT under libstdc++ 13) and has no
    // result to reproduce
    return "row order 1 (-r 1, descending degree) is not provided: the reference sorts with a >= comparator";
  } else if (mode == -1) {
    std::sort(order, order + rows, [&](int a, int b) { return len(a) < len(b); });
  } else if (mode == 2) {
    if (seed != 0) std::srand(seed);
    // std::random_shuffle (removed in C++17) as libstdc++ implements it: element i swaps with rand() % (i + 1)
    for (int64_t i = 1; i < rows; ++i) {
      const int64_t j = std::rand() % (i + 1);
      if (i != j) std::swap(order[i], order[j]);
    }
  } else if (mode != 0) {
    return "row order mode must be 0, 1, -1 or 2";
  }
  return "";
}

int64_t host_partition(const int64_t* grouping, int64_t n, int64_t* part) {
  std::vector<int64_t> g(grouping, grouping + n);
  std::sort(g.begin(), g.end());
  int64_t count = 0;
  for (int64_t i = 0; i < n; ++i)
    if (i == 0 || g[i] != g[i - 1]) part[count++] = i;
  part[count++] = n;
  return count;
}

template <class F>
static void parallel_for(int64_t n, int threads, F fn) {
  if (threads <= 1 || n < 2) {
    fn(0, n);
    return;
  }
  threads = static_cast<int>(std::min<int64_t>(threads, n));
  std::vector<std::thread> pool;
  const int64_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
    if (lo >= hi) break;
    pool.emplace_back([=] { fn(lo, hi); });
  }
  for (auto& th : pool) th.join();
}

// fn(lo, hi) over [0, n) cut into one range per thread of about equal WEIGHT (prefix[i] = weight of the
// items before i): the block-rows of a clustered matrix are sorted dense first, equal counts are not equal work.
template <class F>
static void parallel_for_weighted(int64_t n, int threads, const std::vector<int64_t>& prefix, F fn) {
  if (threads <= 1 || n < 2) {
    fn(0, n);
    return;
  }
  threads = static_cast<int>(std::min<int64_t>(threads, n));
  std::vector<int64_t> cut(threads + 1, 0);
  for (int t = 1; t < threads; ++t) {
    const int64_t target = prefix[n] * t / threads;
    cut[t] = std::max<int64_t>(cut[t - 1], std::lower_bound(prefix.begin(), prefix.begin() + n + 1, target) - prefix.begin());
    cut[t] = std::min(cut[t], n);
  }
  cut[threads] = n;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    if (cut[t] < cut[t + 1]) pool.emplace_back([=] { fn(cut[t], cut[t + 1]); });
  for (auto& th : pool) th.join();
}

// dense: the reference's arrays, mab included.  sparse != nullptr: mab is NOT built; the nonzeros go to
// sparse->nz_* as (offset into the virtual mab, value) and the permutation is kept.
static const char* vbr_fill_impl(int64_t rows_in, int64_t cols_in, const int64_t* rowptr,
                                 const int64_t* colind, const float* val, bool pattern_only,
                                 const int64_t* grouping, int64_t w, int64_t row_block_size,
                                 bool force_fixed, int threads, HostVBR* out, HostVBRSparse* sparse) {
  if (w <= 0) return "column block size must be positive";
  if (force_fixed && row_block_size <= 0) return "force_fixed needs a positive row block size";
  const bool timing = getenv("SPARTA_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(now() - a).count(); };
  const auto tp0 = now();
  std::vector<int64_t> perm(rows_in), part(rows_in + 1);
  {
    // input check (all threads), permutation and partition side by side
    std::vector<char> bad(static_cast<size_t>(std::max(threads, 1)), 0);
    std::thread t_perm([&] { host_permutation(grouping, rows_in, perm.data()); });
    std::thread t_part([&] { part.resize(host_partition(grouping, rows_in, part.data())); });
    for (int64_t i = 0; i < rows_in; ++i)
      if (rowptr[i + 1] < rowptr[i]) bad[0] = 1;
    if (!bad[0]) {
      const int64_t nnz = rowptr[rows_in] - rowptr[0];
      const int T = std::max(threads, 1);
      std::vector<std::thread> pool;
      for (int t = 0; t < T; ++t)
        pool.emplace_back([&, t] {
          const int64_t lo = rowptr[0] + nnz * t / T, hi = rowptr[0] + nnz * (t + 1) / T;
          char b = 0;
          for (int64_t p = lo; p < hi; ++p) b |= (colind[p] < 0) | (colind[p] >= cols_in);
          bad[t] = b ? 2 : 0;
        });
      for (auto& th : pool) th.join();
    }
    t_perm.join();
    t_part.join();
    if (bad[0] == 1) return "rowptr must be non-decreasing";
    for (char b : bad) if (b) return "column index out of range";
  }

  const double t_perm = since(tp0);
  int64_t rows = rows_in, cols = cols_in;
  if (force_fixed) {  // vbr.cpp:142-147: pad to whole blocks, the last block-row absorbs the padding rows
    rows = ((rows_in - 1) / row_block_size + 1) * row_block_size;
    cols = ((cols_in - 1) / w + 1) * w;
    part.back() = rows;
  }
  const int64_t block_rows = static_cast<int64_t>(part.size()) - 1;
  const int64_t block_cols = (cols - 1) / w + 1;
  out->rows = rows; out->cols = cols; out->block_rows = block_rows; out->block_cols = block_cols;
  out->block_col_size = w;
  out->row_part = part;
  out->nzcount.assign(block_rows, 0);

  // nonzeros per block-row: the weight the two passes are balanced on
  std::vector<int64_t> nnz_prefix(block_rows + 1, 0);
  for (int64_t ib = 0; ib < block_rows; ++ib) {
    int64_t cnt = 0;
    for (int64_t r = part[ib]; r < part[ib + 1] && r < rows_in; ++r) cnt += rowptr[perm[r] + 1] - rowptr[perm[r]];
    nnz_prefix[ib + 1] = nnz_prefix[ib] + cnt + 1;
  }
  // pass 1: distinct column blocks per block-row (stamp array per thread)
  std::vector<std::vector<int64_t>> row_jab(block_rows);
  parallel_for_weighted(block_rows, threads, nnz_prefix, [&](int64_t lo, int64_t hi) {
    std::vector<int64_t> stamp(block_cols, -1);
    for (int64_t ib = lo; ib < hi; ++ib) {
      std::vector<int64_t>& list = row_jab[ib];
      for (int64_t r = part[ib]; r < part[ib + 1]; ++r) {
        if (r >= rows_in) break;  // padding rows hold nothing
        const int64_t i = perm[r];
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
          const int64_t jb = colind[p] / w;
          if (stamp[jb] != ib) { stamp[jb] = ib; list.push_back(jb); }
        }
      }
      std::sort(list.begin(), list.end());
      out->nzcount[ib] = static_cast<int64_t>(list.size());
    }
  });

  const double t_pass1 = since(tp0);
  // offsets
  std::vector<int64_t> jab_off(block_rows + 1, 0), mab_off(block_rows + 1, 0);
  for (int64_t ib = 0; ib < block_rows; ++ib) {
    const int64_t h = part[ib + 1] - part[ib];
    jab_off[ib + 1] = jab_off[ib] + out->nzcount[ib];
    mab_off[ib + 1] = mab_off[ib] + out->nzcount[ib] * h * w;
  }
  out->jab.resize(jab_off[block_rows]);
  out->nztot = mab_off[block_rows];
  if (sparse) {
    sparse->nz_ptr.assign(block_rows + 1, 0);
    for (int64_t ib = 0; ib < block_rows; ++ib) {
      int64_t cnt = 0;
      for (int64_t r = part[ib]; r < part[ib + 1] && r < rows_in; ++r) cnt += rowptr[perm[r] + 1] - rowptr[perm[r]];
      sparse->nz_ptr[ib + 1] = sparse->nz_ptr[ib] + cnt;
    }
    if (!sparse->nz_off.alloc(static_cast<size_t>(sparse->nz_ptr[block_rows])) ||
        !sparse->nz_val.alloc(static_cast<size_t>(sparse->nz_ptr[block_rows])))
      return "out of memory";
  } else {
    out->mab.assign(static_cast<size_t>(out->nztot), 0.0f);
  }

  // pass 2: scatter values; block (ib, slot) is column-major with ld = h (vbr.cpp:224)
  parallel_for_weighted(block_rows, threads, nnz_prefix, [&](int64_t lo, int64_t hi) {
    std::vector<int64_t> slot(block_cols, 0);
    for (int64_t ib = lo; ib < hi; ++ib) {
      const std::vector<int64_t>& list = row_jab[ib];
      std::copy(list.begin(), list.end(), out->jab.begin() + jab_off[ib]);
      for (size_t s = 0; s < list.size(); ++s) slot[list[s]] = static_cast<int64_t>(s);
      const int64_t h = part[ib + 1] - part[ib];
      float* base = sparse ? nullptr : out->mab.data() + mab_off[ib];
      int64_t at = sparse ? sparse->nz_ptr[ib] : 0;
      for (int64_t r = part[ib]; r < part[ib + 1]; ++r) {
        if (r >= rows_in) break;
        const int64_t i = perm[r];
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) {
          const int64_t j = colind[p];
          const int64_t off = slot[j / w] * w * h + h * (j % w) + (r - part[ib]);
          const float x = pattern_only ? 1.0f : val[p];
          if (sparse) {
            sparse->nz_off[at] = mab_off[ib] + off;
            sparse->nz_val[at] = x;
            ++at;
          } else {
            base[off] = x;
          }
        }
      }
    }
  });
  if (sparse) sparse->perm.swap(perm);
  if (timing)
    fprintf(stderr, "sparta vbr fill (%s, %d threads): validate + permutation %.1f ms, column-block lists %.1f, values %.1f\n",
            sparse ? "index + nonzero offsets" : "dense mab", threads, t_perm, t_pass1 - t_perm, since(tp0) - t_pass1);
  return "";
}

const char* host_vbr_fill(int64_t rows_in, int64_t cols_in, const int64_t* rowptr,
                          const int64_t* colind, const float* val, bool pattern_only,
                          const int64_t* grouping, int64_t w, int64_t row_block_size,
                          bool force_fixed, int threads, HostVBR* out) {
  return vbr_fill_impl(rows_in, cols_in, rowptr, colind, val, pattern_only, grouping, w, row_block_size, force_fixed,
                       threads, out, nullptr);
}

const char* host_vbr_fill_sparse(int64_t rows_in, int64_t cols_in, const int64_t* rowptr, const int64_t* colind,
                                 const float* val, bool pattern_only, const int64_t* grouping, int64_t w,
                                 int64_t row_block_size, bool force_fixed, int threads, HostVBRSparse* out) {
  return vbr_fill_impl(rows_in, cols_in, rowptr, colind, val, pattern_only, grouping, w, row_block_size, force_fixed,
                       threads, &out->index, out);
}

const char* host_bellpack_from_vbr(int64_t rows, int64_t cols, int64_t bs, const int64_t* nzcount,
                                   const int64_t* jab, const float* mab, int threads,
                                   HostBell* out) {
  if (bs <= 0 || rows % bs || cols % bs) return "rows and cols must be multiples of the block size";
  const int64_t ind_rows = rows / bs;
  int64_t width = 0;
  for (int64_t i = 0; i < ind_rows; ++i) width = std::max(width, nzcount[i]);
  out->blocksize = bs; out->ind_rows = ind_rows; out->ind_cols = width;
  out->col_ind.assign(static_cast<size_t>(ind_rows * width), -1);   // -1 = padding block (:1693)
  out->values.assign(static_cast<size_t>(rows * width * bs), 0.0f);
  std::vector<int64_t> joff(ind_rows + 1, 0);
  for (int64_t i = 0; i < ind_rows; ++i) joff[i + 1] = joff[i] + nzcount[i];
  const int64_t val_cols = width * bs;
  parallel_for(ind_rows, threads, [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      const float* blocks = mab + joff[i] * bs * bs;
      for (int64_t s = 0; s < nzcount[i]; ++s) {
        out->col_ind[i * width + s] = jab[joff[i] + s];
        const float* blk = blocks + s * bs * bs;  // column-major bs x bs
        for (int64_t k = 0; k < bs; ++k)
          for (int64_t r = 0; r < bs; ++r)
            out->values[(i * bs + r) * val_cols + s * bs + k] = blk[k * bs + r];
      }
    }
  });
  return "";
}

// ---- grouping cache (see host_formats.h) ---------------------------------------------------------

static inline uint64_t fnv_mix(uint64_t h, const void* data, size_t bytes) {
  // FNV-1a over 8-byte words (the arrays are int64); the tail bytes one by one
  const uint8_t* p = static_cast<const uint8_t*>(data);
  size_t i = 0;
  for (; i + 8 <= bytes; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 0x100000001B3ull;
  }
  for (; i < bytes; ++i) h = (h ^ p[i]) * 0x100000001B3ull;
  return h;
}

uint64_t blocking_key(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind, int32_t algo,
                      float tau, int64_t block_col_size, int64_t row_block_size, int32_t sim_measure,
                      int32_t use_pattern, int32_t use_groups, int32_t force_fixed_size) {
  uint64_t h = 0xCBF29CE484222325ull;
  const int64_t head[9] = {rows, cols, algo, block_col_size, row_block_size, sim_measure, use_pattern != 0,
                           use_groups != 0, force_fixed_size != 0};
  h = fnv_mix(h, head, sizeof(head));
  h = fnv_mix(h, &tau, sizeof(tau));
  h = fnv_mix(h, rowptr, static_cast<size_t>(rows + 1) * sizeof(int64_t));
  h = fnv_mix(h, colind, static_cast<size_t>(rowptr[rows]) * sizeof(int64_t));
  return h ? h : 1;   // 0 means "unchecked"
}

const char* grouping_save(const char* path, int64_t rows, const int64_t* grouping, uint64_t key, const char* note) {
  const std::string tmp = std::string(path) + ".tmp";
  FILE* f = fopen(tmp.c_str(), "w");
  if (!f) return "cannot create the grouping file";
  for (int64_t i = 0; i < rows; ++i) fprintf(f, "%lld\n", static_cast<long long>(grouping[i]));
  if (fclose(f) != 0) return "write error on the grouping file";
  if (rename(tmp.c_str(), path) != 0) return "cannot move the grouping file into place";
  FILE* k = fopen((std::string(path) + ".key").c_str(), "w");
  if (!k) return "cannot create the grouping key file";
  fprintf(k, "%016llx %lld\n%s\n", static_cast<unsigned long long>(key), static_cast<long long>(rows), note ? note : "");
  fclose(k);
  return "";
}

const char* grouping_load(const char* path, int64_t rows, int64_t* grouping, uint64_t key) {
  if (key) {
    FILE* k = fopen((std::string(path) + ".key").c_str(), "r");
    if (!k) return "miss";
    unsigned long long have = 0;
    long long have_rows = -1;
    const int got = fscanf(k, "%llx %lld", &have, &have_rows);
    fclose(k);
    if (got != 2 || have != key || have_rows != rows) return "miss";
  }
  FILE* f = fopen(path, "r");
  if (!f) return "miss";
  // one integer per line, like read_grouping_file (Matrix_Analysis.cpp:10-32)
  char line[64];
  int64_t i = 0;
  while (i < rows && fgets(line, sizeof(line), f)) {
    char* end = nullptr;
    const long long v = strtoll(line, &end, 10);
    if (end == line) { fclose(f); return "the grouping file holds a line that is not a number"; }
    grouping[i++] = v;
  }
  const bool more = fgets(line, sizeof(line), f) != nullptr && line[0] != '\n' && line[0] != 0;
  fclose(f);
  if (i != rows || more) return "the grouping file does not have one entry per row";
  return "";
}

}  // namespace sparta
