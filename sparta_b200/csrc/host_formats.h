// Host-side format builders (see host_formats.cpp).
#pragma once
#include <stdint.h>
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
int64_t host_partition(const int64_t* grouping, int64_t n, int64_t* part);

const char* host_vbr_fill(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                          const float* val, bool pattern_only, const int64_t* grouping, int64_t w,
                          int64_t row_block_size, bool force_fixed, int threads, HostVBR* out);

const char* host_bellpack_from_vbr(int64_t rows, int64_t cols, int64_t bs, const int64_t* nzcount,
                                   const int64_t* jab, const float* mab, int threads, HostBell* out);

}  // namespace sparta
