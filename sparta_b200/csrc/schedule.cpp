// Host tile scheduler (see schedule.h).
//
// What it replaces in the reference: the per-call cursor arrays
// jab_positions / mab_positions and the "nzs-th block of every block-row"
// double loop of cublas_fixed_blocks_multiply (src/cuda/cuda_utilities.cpp:108-182)
// and the per-level pointer arrays of cublas_blockmat_batched (:811-856).
#include "schedule.h"

#include <algorithm>
#include <queue>
#include <utility>

namespace sparta {

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

const char* build_structure(const BlockRows& br, const ScheduleOptions& opt, Structure* out) {
  Structure& st = *out;
  st = Structure();
  if (br.w <= 0) return "column block size must be positive";
  if (opt.seg_rows < 16 || opt.seg_rows > 256 || opt.seg_rows % 16) return "seg_rows must be a multiple of 16 in [16,256]";
  if (opt.acc_cols != 256 && opt.acc_cols != 512) return "acc_cols must be 256 or 512";
  if (opt.seg_rows > opt.acc_cols) return "seg_rows exceeds acc_cols";
  const int esize = prec_esize(opt.precision);
  const int katom = 128 / esize;    // k elements per 128-byte swizzle row
  const int kstep = 32 / esize;     // k elements per MMA (K = 16 for 16-bit, 8 for tf32)
  const int kalign = 16 / esize;    // TMA needs a 16-byte aligned start along k

  // 1. segments, tagged with their block-row and row offset inside it
  struct SegSrc { int64_t b; int64_t row_off; };
  std::vector<SegSrc> seg_src;
  const int64_t nb = br.count();
  for (int64_t b = 0; b < nb; ++b) {
    const int64_t H = br.height[b];
    const int64_t nblk = br.ptr[b + 1] - br.ptr[b];
    st.n_blocks += nblk;  // zero-height block-rows keep their (empty) blocks in the count
    if (H <= 0) continue;
    st.rows = std::max<int64_t>(st.rows, br.row0[b] + H);
    if (br.row0[b] + H > INT32_MAX) return "shard has more than 2^31 rows";
    st.nztot += nblk * H * br.w;
    for (int64_t off = 0; off < H; off += opt.seg_rows) {
      Segment sg;
      sg.h = static_cast<int32_t>(std::min<int64_t>(opt.seg_rows, H - off));
      sg.h_pad = round_up(sg.h, 16);
      sg.c_row0 = static_cast<int32_t>(br.row0[b] + off);
      sg.tmem_col = 0;
      st.segs.push_back(sg);
      seg_src.push_back({b, off});
    }
  }

  // 2. super-rows: consecutive segments packed into one accumulator stage
  const size_t nseg = st.segs.size();
  size_t s0 = 0;
  std::vector<std::pair<int64_t, int>> merged;  // (jb, member)
  while (s0 < nseg) {
    SuperRow sr{};
    sr.seg_begin = static_cast<int32_t>(s0);
    int cols = 0;
    size_t s1 = s0;
    while (s1 < nseg && (s1 - s0) < static_cast<size_t>(kMaxMembers) &&
           cols + st.segs[s1].h_pad <= opt.acc_cols) {
      st.segs[s1].tmem_col = cols;
      cols += st.segs[s1].h_pad;
      ++s1;
    }
    sr.seg_count = static_cast<int32_t>(s1 - s0);
    sr.n_cols = cols;
    sr.chunk_begin = static_cast<int32_t>(st.chunks.size());

    // 3. merged column-block list of the members
    merged.clear();
    for (size_t s = s0; s < s1; ++s) {
      const int64_t b = seg_src[s].b;
      for (int64_t q = br.ptr[b]; q < br.ptr[b + 1]; ++q)
        merged.emplace_back(br.col[q], static_cast<int>(s - s0));
    }
    std::sort(merged.begin(), merged.end());
    double cost = 12.0 * cols;  // epilogue drain
    size_t i = 0;
    while (i < merged.size()) {
      const int64_t jb = merged[i].first;
      size_t i1 = i;
      uint32_t mask = 0;
      while (i1 < merged.size() && merged[i1].first == jb) {
        mask |= 1u << merged[i1].second;
        ++i1;
      }
      // K slabs of this column block: start at the 16-byte aligned k at or below jb*w
      const int64_t kblk = jb * br.w;
      const int64_t ka = kblk / kalign * kalign;
      const int shift = static_cast<int>(kblk - ka);
      const int64_t atoms = (shift + br.w + katom - 1) / katom;
      for (int64_t a = 0; a < atoms; ++a) {
        const int64_t k_lo = a * katom - shift;  // block-local k of image column 0
        const int k_used = static_cast<int>(std::min<int64_t>(katom, br.w - k_lo));
        Chunk ch{};
        const int64_t k0 = ka + a * katom;
        if (k0 > INT32_MAX) return "k index exceeds 2^31";
        ch.k0 = static_cast<int32_t>(k0);
        ch.mask = mask;
        ch.ksteps = (k_used + kstep - 1) / kstep;
        if ((st.a_bytes >> 4) > UINT32_MAX) return "packed A exceeds 64 GiB";
        ch.a_off16 = static_cast<uint32_t>(st.a_bytes >> 4);
        uint32_t bytes = 0;
        for (size_t t = i; t < i1; ++t) {
          const int m = merged[t].second;
          const size_t s = s0 + m;
          const int64_t b = seg_src[s].b;
          // position of block jb inside block-row b
          const int64_t* cb = br.col.data() + br.ptr[b];
          const int64_t* ce = br.col.data() + br.ptr[b + 1];
          const int64_t q = br.ptr[b] + (std::lower_bound(cb, ce, jb) - cb);
          PackJob job;
          job.src_rs = br.rs[b];
          job.src_ks = br.ks[b];
          job.src_base = br.src[q] + seg_src[s].row_off * job.src_rs;
          job.h = st.segs[s].h;
          job.h_pad = st.segs[s].h_pad;
          job.k_lo = static_cast<int32_t>(k_lo);
          job.k_w = static_cast<int32_t>(br.w);
          job.pad_[0] = job.pad_[1] = job.pad_[2] = 0;
          job.dst_off16 = static_cast<uint32_t>((st.a_bytes + bytes) >> 4);
          st.jobs.push_back(job);
          bytes += static_cast<uint32_t>(job.h_pad) * 128u;
        }
        ch.a_bytes = bytes;
        st.a_bytes += bytes;
        st.max_chunk_bytes = std::max(st.max_chunk_bytes, bytes);
        st.chunks.push_back(ch);
        const double tensor = ch.ksteps * (bytes / 128.0) * 0.5;  // N/2 cycles per MMA at M=128
        const double memory = (kPanelBytes + bytes) / 48.0;
        cost += std::max(tensor, memory) + 40.0;
      }
      i = i1;
    }
    sr.chunk_count = static_cast<int32_t>(st.chunks.size()) - sr.chunk_begin;
    st.srows.push_back(sr);
    st.srow_cost.push_back(cost);
    s0 = s1;
  }
  if (st.chunks.size() > static_cast<size_t>(INT32_MAX)) return "too many chunks";
  return "";
}

const char* build_assignment(const Structure& st, const ScheduleOptions& opt, int64_t n,
                             Assignment* out) {
  Assignment& as = *out;
  as = Assignment();
  if (n <= 0 || n > INT32_MAX - kTileJ) return "invalid number of B columns";
  const int64_t tiles = (n + kTileJ - 1) / kTileJ;
  const int64_t n_items = static_cast<int64_t>(st.srows.size()) * tiles;
  if (n_items > INT32_MAX) return "too many work items";
  as.items.reserve(n_items);
  std::vector<double> cost;
  cost.reserve(n_items);
  for (size_t s = 0; s < st.srows.size(); ++s)
    for (int64_t t = 0; t < tiles; ++t) {
      as.items.push_back(Item{static_cast<int32_t>(s), static_cast<int32_t>(t * kTileJ)});
      cost.push_back(st.srow_cost[s]);
    }
  as.grid = static_cast<int>(std::min<int64_t>(opt.num_ctas, n_items));
  as.cta_ptr.assign(as.grid + 1, 0);
  if (as.grid == 0) return "";

  // longest-processing-time-first onto the persistent CTAs
  std::vector<int32_t> order(n_items);
  for (int64_t i = 0; i < n_items; ++i) order[i] = static_cast<int32_t>(i);
  std::stable_sort(order.begin(), order.end(),
                   [&](int32_t a, int32_t b) { return cost[a] > cost[b]; });
  typedef std::pair<double, int> Load;
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int c = 0; c < as.grid; ++c) heap.push(Load(0.0, c));
  std::vector<std::vector<int32_t>> per_cta(as.grid);
  for (int32_t id : order) {
    Load l = heap.top();
    heap.pop();
    per_cta[l.second].push_back(id);
    l.first += cost[id];
    heap.push(l);
  }
  double total = 0;
  while (!heap.empty()) {
    as.max_cta_cost = std::max(as.max_cta_cost, heap.top().first);
    total += heap.top().first;
    heap.pop();
  }
  as.mean_cta_cost = total / as.grid;
  as.cta_items.reserve(n_items);
  for (int c = 0; c < as.grid; ++c) {
    as.cta_ptr[c] = static_cast<int32_t>(as.cta_items.size());
    as.cta_items.insert(as.cta_items.end(), per_cta[c].begin(), per_cta[c].end());
  }
  as.cta_ptr[as.grid] = static_cast<int32_t>(as.cta_items.size());
  return "";
}

void partition_block_rows(int64_t block_rows, const int64_t* row_part, const int64_t* nzcount,
                          int parts, int64_t* cuts) {
  std::vector<double> prefix(block_rows + 1, 0.0);
  for (int64_t b = 0; b < block_rows; ++b)
    prefix[b + 1] = prefix[b] + static_cast<double>(nzcount[b]) * (row_part[b + 1] - row_part[b]);
  const double total = prefix[block_rows];
  cuts[0] = 0;
  for (int i = 1; i < parts; ++i) {
    const double target = total * i / parts;
    int64_t b = std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin();
    if (b > 0 && target - prefix[b - 1] < prefix[b] - target) --b;
    b = std::max(b, cuts[i - 1]);
    cuts[i] = std::min(b, block_rows);
  }
  cuts[parts] = block_rows;
}

}  // namespace sparta
