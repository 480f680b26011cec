// Host-side row clustering (see blocking.cpp).
#pragma once
#include <stdint.h>

namespace sparta {

struct BlockingParams {      // named after the reference CLI flags (include/input.h:15-42)
  int     algo = 3;              // -a: 0 iterative, 2 fixed, 3 clocked, 4 queue, 5 max-size (keeper)
  float   tau = 0.1f;            // -t
  int64_t block_col_size = 3;    // -b
  int64_t row_block_size = 3;    // -B
  int     sim_measure = 1;       // -m: 0 Hamming, 1 Jaccard, 2/3 their probe variants
  bool    use_pattern = true;    // -p
  bool    use_groups = false;    // -g
  bool    force_fixed_size = false;  // -F
  bool    force_list_model = false;  // testing: Jaccard through the column-list model
};

struct BlockingStats {       // the counters save_blocking_data prints (utilities.cpp:175-233)
  int64_t comparisons = 0, merges = 0;
  float   average_merge_tau = 0, average_row_distance = 0;
};

// grouping[rows] receives one group id per row.  Returns "" or a static error string.
const char* host_blocking(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                          const BlockingParams& p, int64_t* grouping, BlockingStats* stats);

}  // namespace sparta
