// Host-side format builders (see host_formats.cpp).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <vector>

namespace sparta {

struct HostVBR {   // same fields as the reference's struct VBR (include/matrices.h:93-122)
  int64_t rows = 0, cols = 0, block_rows = 0, block_cols = 0, block_col_size = 0, nztot = 0;
  std::vector<int64_t> row_part, nzcount, jab;
  std::vector<float> mab;
};

struct HostBell {  // the out-params of prepare_cusparse_BLOCKEDELLPACK (cuda_utilities.cpp:1656)
  int64_t blocksize = 0, ind_rows = 0, ind_cols = 0;
  std::vector<int64_t> col_ind;
  std::vector<float> values;
};

void host_permutation(const int64_t* grouping, int64_t n, int64_t* perm);

// The reference's -r row reorderings (include/input.h:30, src/general/csr.cpp:123-166): new row i is old
// row order[i].  mode -1: ascending degree, 2: scramble (mode 1, descending degree, is undefined behaviour in
// the reference -- a >= comparator handed to std::sort -- and is refused).  The scramble is
// std::random_shuffle driven by std::rand, which the reference seeds from -s (input.h:111-114); the same
// glibc generator is seeded and stepped here (seed 0 = leave the generator as it is).
const char* host_row_order(int64_t rows, const int64_t* rowptr, int32_t mode, uint32_t seed, int64_t* order);
int64_t host_partition(const int64_t* grouping, int64_t n, int64_t* part);

const char* host_vbr_fill(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                          const float* val, bool pattern_only, const int64_t* grouping, int64_t w,
                          int64_t row_block_size, bool force_fixed, int threads, HostVBR* out);

// The same build WITHOUT materialising the dense blocks: row_part / nzcount / jab / nztot as above
// (out->mab stays empty), the permutation, and the nonzeros as (element offset into the virtual mab,
// value) pairs -- what the device needs to rebuild the blocks itself (28 MB instead of 4.45 GB at
// BASELINE config #3).  Entries are grouped by block-row: block-row ib owns [nz_ptr[ib], nz_ptr[ib+1]).
template <class T>
struct RawBuf {                  // uninitialised storage: the filling threads touch the pages first
  T* p = nullptr;
  size_t n = 0;
  // optional allocator of the owner (the C ABI layer hands out page-locked blocks so that the arrays can
  // cross PCIe by DMA straight from where the threads wrote them); tag is returned to release_fn
  void* (*acquire_fn)(size_t bytes, bool* tag) = nullptr;
  void (*release_fn)(void* p, bool tag) = nullptr;
  bool tag = false;
  RawBuf() = default;
  RawBuf(const RawBuf&) = delete;
  RawBuf& operator=(const RawBuf&) = delete;
  ~RawBuf() { drop(); }
  void drop() {
    if (p) { if (release_fn) release_fn(p, tag); else free(p); }
    p = nullptr;
  }
  bool alloc(size_t count) {
    drop();
    n = count;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    p = static_cast<T*>(acquire_fn ? acquire_fn(bytes, &tag) : malloc(bytes));
    return p != nullptr;
  }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};
struct HostVBRSparse {
  HostVBR index;                 // mab empty
  std::vector<int64_t> perm;     // blocked row r is original row perm[r] (rows beyond the input: padding)
  std::vector<int64_t> nz_ptr;   // [block_rows + 1]
  RawBuf<int64_t> nz_off;        // element offset of each nonzero inside the virtual mab
  RawBuf<float> nz_val;
};
const char* host_vbr_fill_sparse(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                                 const float* val, bool pattern_only, const int64_t* grouping, int64_t w,
                                 int64_t row_block_size, bool force_fixed, int threads, HostVBRSparse* out);

const char* host_bellpack_from_vbr(int64_t rows, int64_t cols, int64_t bs, const int64_t* nzcount,
                                   const int64_t* jab, const float* mab, int threads, HostBell* out);

// Host threads a call may start: the CPUs this process may run on (affinity mask), capped by the cgroup CPU
// quota of the container (a process that runs more threads than its quota is throttled in 100 ms periods --
// measured as 100 ms stalls inside cudaMemcpyAsync with two ranks on one box), divided by LOCAL_WORLD_SIZE
// when a launcher that starts one process per GPU set it, overridden by SPARTA_THREADS; at most `cap`.
int host_thread_budget(int cap);

// ---- grouping cache -----------------------------------------------------------------------------
// The reference can persist a grouping as `<outfile>.g`, one group id per line
// (test/general/Matrix_Blocking.cpp:24-32, src/general/utilities.cpp:240-243), and reload it
// (test/general/Matrix_Analysis.cpp:10-32).  The cache keeps exactly that file format, so either
// side can read the other's files, and adds a sidecar `<file>.key` holding a 64-bit key of the
// CSR pattern and the blocking flags, so a stale file is never served for a different input.
uint64_t blocking_key(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind, int32_t algo,
                      float tau, int64_t block_col_size, int64_t row_block_size, int32_t sim_measure,
                      int32_t use_pattern, int32_t use_groups, int32_t force_fixed_size);
const char* grouping_save(const char* path, int64_t rows, const int64_t* grouping, uint64_t key, const char* note);
// key == 0: do not check the sidecar.  Returns "" on success; "miss" when the file or its key is absent / different.
const char* grouping_load(const char* path, int64_t rows, int64_t* grouping, uint64_t key);

}  // namespace sparta
