// Row clustering of the product's host layer: grouping = f(CSR, tau, w, ...), the
// input of the VBR build.  Output must equal the reference's
// BlockingEngine::GetGrouping (src/general/blocking.cpp:633-676) element for
// element, because the north star asks for bit-identical VBR index arrays.
//
// The reference evaluates every row-vs-pattern distance by walking two sorted
// COLUMN lists (blocking.cpp:923-994) and re-materialises the pattern as a
// column list on every merge (utilities.cpp:145-173).  For the Jaccard measure
// (the default, -m 1) only column-BLOCK membership matters, so the fast path
// here keeps the pattern as a bitmap over column blocks plus the pattern's
// largest column (needed to reproduce merge_rows' truncation, see
// BlockPattern::merge) and a distinct-block list per row: a distance is a few
// dozen bit tests instead of a merge walk over thousands of entries.  The other
// measures go through ListPattern, a plain column-list model.
//
// What cannot be changed without changing the output and is therefore kept:
//   * the sequential seed / scan order and the `distances[]` pruning array whose
//     element 0 starts at -1 and all others at 0 (blocking.cpp:159,255,436);
//   * `-a 5`'s candidate set: a std::set<pair<float,intT>> trimmed by advancing
//     end() (blocking.cpp:503-511) -- undefined behaviour whose libstdc++
//     outcome is reproduced by issuing the same std::set calls;
//   * fp32 accumulation order of the merge statistics.
#include "blocking.h"

// `-a 5` reproduces an undefined-behaviour outcome of the reference (advancing std::set::end(),
// src/general/blocking.cpp:509-511) by issuing the same calls against the same standard library; with
// another std::set implementation the grouping would silently differ, so refuse to build there.
#include <set>
#ifndef __GLIBCXX__
#error "sparta_b200/csrc/blocking.cpp needs libstdc++: -a 5 replays the reference's std::set calls (see run_keeper)"
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <thread>
#include <numeric>
#include <set>
#include <utility>
#include <vector>

#include "host_formats.h"

namespace sparta {

namespace {

struct Rows {
  int64_t rows, cols, w, block_cols;
  const int64_t* rowptr;
  const int64_t* colind;
  std::vector<int64_t> bptr;   // distinct column blocks per row (CSR over rows)
  std::vector<int32_t> bidx;
  int64_t len(int64_t i) const { return rowptr[i + 1] - rowptr[i]; }
  const int64_t* row(int64_t i) const { return colind + rowptr[i]; }
};

// ---- pattern models --------------------------------------------------------

// Jaccard on column blocks (JaccardDistanceGroup, blocking.cpp:923-994, count_zeros = 1).
class BlockPattern {
 public:
  explicit BlockPattern(const Rows& m) : m_(m), bits_((m.block_cols + 63) / 64, 0) {}

  void seed(int64_t i) {
    std::fill(bits_.begin(), bits_.end(), 0);
    nblocks_ = 0;
    const int64_t n = m_.len(i);
    max_col_ = n ? m_.row(i)[n - 1] : -1;
    for (int64_t p = m_.bptr[i]; p < m_.bptr[i + 1]; ++p) set(m_.bidx[p]);
  }

  float dist(int64_t j, int64_t gsize) const {
    const int64_t nb = m_.bptr[j + 1] - m_.bptr[j];
    if (max_col_ < 0 && nb == 0) return 0;
    if (max_col_ < 0 || nb == 0) return 1;
    int64_t inter = 0;
    for (int64_t p = m_.bptr[j]; p < m_.bptr[j + 1]; ++p) {
      const int32_t b = m_.bidx[p];
      inter += (bits_[b >> 6] >> (b & 63)) & 1;
    }
    const int64_t only_pattern = nblocks_ - inter, only_row = nb - inter;
    const int64_t count = only_pattern + only_row * gsize;
    return (2.0 * count) / (nblocks_ * gsize + nb + count);   // double, narrowed on return (:993)
  }

  // merge_rows (utilities.cpp:145-173) is not a set union: with c* = the largest entry
  // of the row that is <= the pattern's largest column, pattern entries above c* are
  // dropped (all of them when no such entry exists), then the whole row is added.  On
  // column blocks that is: clear every block above block(c*), OR in the row's blocks.
  void merge(int64_t j) {
    const int64_t n = m_.len(j);
    const int64_t* r = m_.row(j);
    if (n == 0) {  // empty row: the result is empty
      std::fill(bits_.begin(), bits_.end(), 0);
      nblocks_ = 0;
      max_col_ = -1;
      return;
    }
    const int64_t* up = std::upper_bound(r, r + n, max_col_);
    if (up == r) {
      std::fill(bits_.begin(), bits_.end(), 0);
      nblocks_ = 0;
    } else {
      clear_above(up[-1] / m_.w);
    }
    for (int64_t p = m_.bptr[j]; p < m_.bptr[j + 1]; ++p) set(m_.bidx[p]);
    max_col_ = r[n - 1];
  }

 private:
  void set(int32_t b) {
    uint64_t& word = bits_[b >> 6];
    const uint64_t bit = 1ull << (b & 63);
    nblocks_ += !(word & bit);
    word |= bit;
  }
  void clear_above(int64_t b) {   // keep blocks <= b
    const size_t wi = static_cast<size_t>(b >> 6);
    const int sh = static_cast<int>(b & 63);
    if (sh != 63) {
      const uint64_t keep = (2ull << sh) - 1;
      nblocks_ -= __builtin_popcountll(bits_[wi] & ~keep);
      bits_[wi] &= keep;
    }
    for (size_t k = wi + 1; k < bits_.size(); ++k) {
      nblocks_ -= __builtin_popcountll(bits_[k]);
      bits_[k] = 0;
    }
  }
  const Rows& m_;
  std::vector<uint64_t> bits_;
  int64_t nblocks_ = 0;
  int64_t max_col_ = -1;
};

// Column-list model for the remaining measures: 0 HammingDistanceGroup (:859-921),
// 2/3 the "OPENMP" variants (:720-856; probe each distinct block of the row with a
// lower_bound into the pattern, count_zeros = 0), and 1 as a cross-check of BlockPattern.
class ListPattern {
 public:
  ListPattern(const Rows& m, int measure) : m_(m), measure_(measure) {}
  void seed(int64_t i) { pat_.assign(m_.row(i), m_.row(i) + m_.len(i)); }

  float dist(int64_t j, int64_t gsize) const {
    const int64_t na = static_cast<int64_t>(pat_.size()), nb = m_.len(j);
    const bool jaccard = measure_ == 1 || measure_ == 3;
    if (na == 0 && nb == 0) return 0;
    if (na == 0 || nb == 0) return jaccard ? 1.0f : static_cast<float>(std::max(na * gsize, nb));
    const int64_t* r = m_.row(j);
    int64_t blocks_a = 0, blocks_b = 0, only_a = 0, only_b = 0;
    if (measure_ <= 1) {
      int64_t i = 0, k = 0;
      while (i < na || k < nb) {
        const int64_t a = i < na ? pat_[i] / m_.w : INT64_MAX, b = k < nb ? r[k] / m_.w : INT64_MAX;
        const int64_t cur = std::min(a, b);
        if (a == cur) { ++blocks_a; while (i < na && pat_[i] / m_.w == cur) ++i; }
        if (b == cur) { ++blocks_b; while (k < nb && r[k] / m_.w == cur) ++k; }
        only_a += (a == cur && b != cur);
        only_b += (b == cur && a != cur);
      }
      const int64_t count = only_a + only_b * gsize;
      if (!jaccard) return static_cast<float>(count);
      return (2.0 * count) / (blocks_a * gsize + blocks_b + count);
    }
    int64_t last = -1;
    for (int64_t i = 0; i < na; ++i)
      if (pat_[i] / m_.w != last) { last = pat_[i] / m_.w; ++blocks_a; }
    int64_t inter = 0, diff = 0;
    last = -1;
    for (int64_t k = 0; k < nb; ++k) {
      const int64_t b = r[k] / m_.w;
      if (b == last) continue;
      last = b;
      std::vector<int64_t>::const_iterator it = std::lower_bound(pat_.begin(), pat_.end(), b * m_.w);
      // a probe that runs off the end counts as "different" in both variants (:774, :841)
      if (it != pat_.end() && *it / m_.w == b) ++inter; else ++diff;
    }
    const int64_t count = diff + (blocks_a - inter) * gsize;
    if (!jaccard) return static_cast<float>(count);
    return (2.0 * count) / (blocks_a * gsize + (diff + inter) + count);
  }

  void merge(int64_t j) {
    const int64_t n = m_.len(j);
    const int64_t* r = m_.row(j);
    std::vector<int64_t> out;
    if (n && !pat_.empty()) {
      const int64_t* up = std::upper_bound(r, r + n, pat_.back());
      if (up != r) {
        // pattern entries <= the last row entry that still lies inside the pattern's range
        std::vector<int64_t>::const_iterator cut = std::upper_bound(pat_.begin(), pat_.end(), up[-1]);
        out.resize((cut - pat_.begin()) + n);
        out.erase(std::set_union(pat_.cbegin(), cut, r, r + n, out.begin()), out.end());
        pat_.swap(out);
        return;
      }
    }
    pat_.assign(r, r + n);
  }

 private:
  const Rows& m_;
  int measure_;
  std::vector<int64_t> pat_;
};

// `float distances[rows] = {-1}`: element 0 is -1, every other element 0.
std::vector<float> initial_distances(int64_t rows) {
  std::vector<float> d(static_cast<size_t>(std::max<int64_t>(rows, 1)), 0.0f);
  d[0] = -1;
  return d;
}

struct Acc {
  int64_t comparisons = 0, merges = 0;
  float merge_tau = 0, row_distance = 0;
};

inline bool pruned(std::vector<float>& d, int64_t i, int64_t j, float tau) {
  if (d[i] != -1 && d[j] != -1 && std::abs(d[i] - d[j]) > tau) {
    d[j] = -1;
    return true;
  }
  return false;
}

// ---- speculative evaluation of a scan window on all host threads -----------------------------
// The scans of -a 3 / -a 4 are sequential in their DECISIONS (a merge changes the pattern every later
// distance is taken against) but merges are rare: one per ~20 000 comparisons on the R-MAT of BASELINE
// config #4.  A window of upcoming candidates is therefore evaluated against the CURRENT pattern by all
// threads (pure reads), then committed in order by one thread exactly like the sequential loop; at the
// first merge the rest of the window is thrown away and re-evaluated against the new pattern.  Same
// distances (same function, same operands), same order of decisions, same counters: the grouping is
// identical by construction, and tests/test_blocking.py checks it against the reference build.
class ScanPool {
 public:
  explicit ScanPool(int threads) : n_(std::max(1, threads)) {
    for (int t = 1; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
  }
  ~ScanPool() {
    stop_.store(true, std::memory_order_relaxed);
    gen_.fetch_add(1, std::memory_order_release);
    for (std::thread& w : workers_) w.join();
  }
  int threads() const { return n_; }
  // fn(lo, hi) over contiguous slices of [0, count); returns when all slices are done
  template <class F>
  void run(int64_t count, const F& fn) {
    if (n_ == 1 || count < 4 * n_) { fn(0, count); return; }
    ctx_ = &fn;
    call_ = [](const void* c, int64_t lo, int64_t hi) { (*static_cast<const F*>(c))(lo, hi); };
    count_ = count;
    pending_.store(n_ - 1, std::memory_order_relaxed);
    gen_.fetch_add(1, std::memory_order_release);
    fn(0, count / n_);
    for (int spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins)
      if (spins > 2000) std::this_thread::yield();
  }

 private:
  void loop(int t) {
    uint64_t seen = 0;
    for (;;) {
      int spins = 0;
      while (gen_.load(std::memory_order_acquire) == seen) {
        if (++spins > 4000) {
          if (spins > 20000) std::this_thread::sleep_for(std::chrono::microseconds(50));
          else std::this_thread::yield();
        }
      }
      ++seen;
      if (stop_.load(std::memory_order_relaxed)) return;
      call_(ctx_, count_ * t / n_, count_ * (t + 1) / n_);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  const int n_;
  std::vector<std::thread> workers_;
  std::atomic<uint64_t> gen_{0};
  std::atomic<int> pending_{0};
  std::atomic<bool> stop_{false};
  const void* ctx_ = nullptr;
  void (*call_)(const void*, int64_t, int64_t) = nullptr;
  int64_t count_ = 0;
};

struct Probe {
  uint8_t state;   // 0: row already in a group, 1: pruned by the distances[] rule, 2: distance evaluated
  float dist;
};
inline bool would_prune(const std::vector<float>& d, int64_t i, int64_t j, float tau) {
  return d[i] != -1 && d[j] != -1 && std::abs(d[i] - d[j]) > tau;
}
// probes[t] for the candidates row_of(t), t in [0, count): pruning is tested BEFORE group membership,
// like the sequential loops do (a pruned row has its distance reset even when it is already taken)
template <class P, class RowOf>
void probe_window(ScanPool& pool, const P& pat, int64_t gsize, int64_t i, float tau, const std::vector<float>& d,
                  const int64_t* g, const RowOf& row_of, int64_t count, Probe* out) {
  pool.run(count, [&](int64_t lo, int64_t hi) {
    for (int64_t t = lo; t < hi; ++t) {
      const int64_t j = row_of(t);
      if (would_prune(d, i, j, tau)) { out[t].state = 1; continue; }
      if (g[j] != -1) { out[t].state = 0; continue; }
      out[t].state = 2;
      out[t].dist = pat.dist(j, gsize);
    }
  });
}
constexpr int64_t kWindowMin = 2048, kWindowMax = 65536;
// Half of the thread budget, at most 8: the windows are short (tens of microseconds of work), and with
// every core spinning on the hand-over the scan got SLOWER than sequential on an 8-CPU container
// (R-MAT 2^16, -a 4: 1 thread 50.1 s, 4 threads 20.1 s, 8 threads 55.3 s).
inline int scan_threads() { return std::max(1, std::min(8, host_thread_budget(16) / 2)); }

// -a 0, IterativeBlockingPattern (blocking.cpp:89-154): strict `<`; the pattern merge
// runs whatever use_pattern says (the `if` at :128 guards only a timer macro).
template <class P>
void run_iterative(const Rows& m, P& pat, float tau, bool use_size, int64_t* g, Acc& st) {
  for (int64_t i = 0; i < m.rows; ++i) {
    if (g[i] != -1) continue;
    pat.seed(i);
    int64_t gsize = 1;
    g[i] = i;
    for (int64_t j = i + 1; j < m.rows; ++j) {
      if (g[j] != -1) continue;
      ++st.comparisons;
      if (pat.dist(j, gsize) < tau) {
        ++st.merges;
        g[j] = i;
        pat.merge(j);
        if (use_size) ++gsize;
      }
    }
  }
}

// -a 3, IterativeBlockingPatternCLOCKED (blocking.cpp:156-243)
template <class P>
void run_clocked(const Rows& m, P& pat, float tau, bool use_size, bool use_pattern, int64_t* g, Acc& st) {
  std::vector<float> d = initial_distances(m.rows);
  ScanPool pool(scan_threads());
  std::vector<Probe> probes(static_cast<size_t>(kWindowMax));
  for (int64_t i = 0; i < m.rows; ++i) {
    if (g[i] != -1) continue;
    pat.seed(i);
    int64_t gsize = 1;
    g[i] = i;
    int64_t window = kWindowMin;
    for (int64_t j0 = i + 1; j0 < m.rows;) {
      const int64_t count = std::min(window, m.rows - j0);
      probe_window(pool, pat, gsize, i, tau, d, g, [j0](int64_t t) { return j0 + t; }, count, probes.data());
      int64_t t = 0;
      bool merged = false;
      for (; t < count && !merged; ++t) {
        const int64_t j = j0 + t;
        if (probes[t].state == 1) { d[j] = -1; continue; }
        if (probes[t].state == 0) continue;
        ++st.comparisons;
        const float dist = probes[t].dist;
        d[j] = dist;
        if (dist <= tau) {
          st.merge_tau += dist;
          st.row_distance += j - i;
          ++st.merges;
          g[j] = i;
          if (use_pattern) pat.merge(j);
          if (use_size) ++gsize;
          merged = use_pattern || use_size;   // later distances of the window are stale only if the pattern or the size moved
        }
      }
      j0 += t;
      window = merged ? kWindowMin : std::min(window * 2, kWindowMax);
    }
  }
}

// -a 4, IterativeBlockingQueue (blocking.cpp:245-338): rejected rows are re-queued in order.
template <class P>
void run_queue(const Rows& m, P& pat, float tau, bool use_size, bool use_pattern, int64_t* g, Acc& st) {
  std::vector<float> d = initial_distances(m.rows);
  std::vector<int64_t> pending(m.rows), kept;
  std::iota(pending.begin(), pending.end(), static_cast<int64_t>(0));
  kept.reserve(m.rows);
  ScanPool pool(scan_threads());
  std::vector<Probe> probes(static_cast<size_t>(kWindowMax));
  while (!pending.empty()) {
    const int64_t i = pending[0];
    pat.seed(i);
    int64_t gsize = 1;
    g[i] = i;
    kept.clear();
    int64_t window = kWindowMin;
    const int64_t n_pending = static_cast<int64_t>(pending.size());
    for (int64_t q0 = 1; q0 < n_pending;) {
      const int64_t count = std::min(window, n_pending - q0);
      const int64_t* cand = pending.data() + q0;
      // (every pending row is still without a group: g[j] == -1, state 0 cannot occur)
      probe_window(pool, pat, gsize, i, tau, d, g, [cand](int64_t t) { return cand[t]; }, count, probes.data());
      int64_t t = 0;
      bool merged = false;
      for (; t < count && !merged; ++t) {
        const int64_t j = cand[t];
        if (probes[t].state == 1) { d[j] = -1; kept.push_back(j); continue; }
        ++st.comparisons;
        const float dist = probes[t].dist;
        d[j] = dist;
        if (dist > tau) { kept.push_back(j); continue; }
        st.merge_tau += dist;
        st.row_distance += j - i;
        ++st.merges;
        g[j] = i;
        if (use_pattern) pat.merge(j);
        if (use_size) ++gsize;
        merged = use_pattern || use_size;
      }
      q0 += t;
      window = merged ? kWindowMin : std::min(window * 2, kWindowMax);
    }
    pending.swap(kept);
  }
}

// -a 5 runs IterativeBlockingKeeper (blocking.cpp:433-549, dispatched at :655).
// Node storage of the candidate set of -a 5: one size class, a free list over slabs; the tree's shape and
// every decision taken on it are those of std::set with std::allocator (the allocator only supplies memory).
struct NodeArena {
  size_t node_bytes = 0;
  void* free_list = nullptr;
  std::vector<void*> slabs;
  ~NodeArena() { for (void* p : slabs) ::operator delete(p); }
  void* take(size_t bytes) {
    if (node_bytes == 0) node_bytes = (bytes + 15) / 16 * 16;
    if (bytes > node_bytes) return nullptr;
    if (!free_list) {
      const size_t count = 4096;
      char* slab = static_cast<char*>(::operator new(node_bytes * count));
      slabs.push_back(slab);
      for (size_t i = 0; i < count; ++i) {
        void* p = slab + i * node_bytes;
        *static_cast<void**>(p) = free_list;
        free_list = p;
      }
    }
    void* p = free_list;
    free_list = *static_cast<void**>(p);
    return p;
  }
  void give(void* p) {
    *static_cast<void**>(p) = free_list;
    free_list = p;
  }
};
template <class T>
struct PoolAllocator {
  typedef T value_type;
  NodeArena* arena;
  explicit PoolAllocator(NodeArena* a) : arena(a) {}
  template <class U> PoolAllocator(const PoolAllocator<U>& o) : arena(o.arena) {}
  T* allocate(size_t n) {
    void* p = n == 1 ? arena->take(sizeof(T)) : nullptr;
    return static_cast<T*>(p ? p : ::operator new(n * sizeof(T)));
  }
  void deallocate(T* p, size_t n) {
    if (n == 1 && sizeof(T) <= arena->node_bytes) arena->give(p); else ::operator delete(p);
  }
  template <class U> bool operator==(const PoolAllocator<U>& o) const { return arena == o.arena; }
  template <class U> bool operator!=(const PoolAllocator<U>& o) const { return arena != o.arena; }
};

// What `it = s.end(); std::advance(it, k);` yields in libstdc++ for a non-empty set and k >= 1, in O(1).
// Incrementing the header node (end()) goes to its right link -- the LARGEST element R -- and from there
// down R's left links (_Rb_tree_increment); R has no right child, so by the red-black invariants its left
// subtree is empty or one red leaf L.  From L the successor is R, from R it is the header again.  k steps
// from end() therefore walk the cycle [L,] R, header, [L,] R, header, ...  (The reference does this walk
// at blocking.cpp:509-511; it is undefined behaviour there and here it is reproduced, not repaired.
// std::advance itself made 64 dependent pointer loads per rejected row: 94 % of the -a 5 time.)
template <class S>
typename S::iterator advance_from_end(S& s, size_t k) {
  const std::_Rb_tree_node_base* hdr = s.end()._M_node;
  const std::_Rb_tree_node_base* R = hdr->_M_right;
  const std::_Rb_tree_node_base* L = R->_M_left;
  const std::_Rb_tree_node_base* cyc[3];
  size_t c = 0;
  if (L) cyc[c++] = L;
  cyc[c++] = R;
  cyc[c++] = hdr;
  return typename S::iterator(cyc[(k - 1) % c]);
}

template <class P>
void run_keeper(const Rows& m, P& pat, float tau, int64_t max_rows, bool use_pattern, int64_t* g, Acc& st) {
  typedef std::pair<float, int64_t> Cand;
  typedef std::set<Cand, std::less<Cand>, PoolAllocator<Cand>> Best;   // same tree algorithms, nodes from a free list
  NodeArena arena;
  std::vector<float> d = initial_distances(m.rows);
  std::vector<int64_t> members;
  for (int64_t i = 0; i < m.rows; ++i) {
    if (g[i] != -1) continue;
    Best best{std::less<Cand>(), PoolAllocator<Cand>(&arena)};
    members.assign(1, i);
    const int64_t gid = i + m.rows;
    pat.seed(i);
    int64_t gsize = 1;
    g[i] = gid;
    for (int64_t j = i + 1; j < m.rows && gsize != max_rows; ++j) {
      if (pruned(d, i, j, tau)) continue;
      if (g[j] != -1) continue;
      ++st.comparisons;
      const float dist = pat.dist(j, gsize);   // the group size is always passed here (:480)
      d[j] = dist;
      if (dist <= tau) {
        st.merge_tau += dist;
        st.row_distance += j - i;
        ++st.merges;
        g[j] = gid;
        members.push_back(j);
        if (use_pattern) pat.merge(j);
        ++gsize;
      } else {
        best.insert(std::make_pair(dist, j));
        if (best.size() > static_cast<size_t>(max_rows) - members.size()) {
          // The reference trims with advance(end(), k); erase(it, end()) (:509-511).  Walking
          // forward from end() is undefined; the same calls are made so that libstdc++ visits
          // the same tree nodes and erases the same elements.
          best.erase(advance_from_end(best, static_cast<size_t>(max_rows) - members.size()), best.end());
        }
      }
    }
    if (gsize < max_rows)
      for (Best::iterator it = best.begin(); it != best.end() && gsize != max_rows; ++it) {
        g[it->second] = gid;
        members.push_back(it->second);
        ++gsize;
      }
    if (gsize == max_rows)   // complete groups are renumbered so they sort first (:527-533)
      for (size_t t = 0; t < members.size(); ++t) g[members[t]] -= m.rows;
  }
}

template <class P>
void dispatch(const Rows& m, P& pat, const BlockingParams& p, int64_t* g, Acc& st) {
  switch (p.algo) {
    case 0: run_iterative(m, pat, p.tau, p.use_groups, g, st); break;
    case 3: run_clocked(m, pat, p.tau, p.use_groups, p.use_pattern, g, st); break;
    case 4: run_queue(m, pat, p.tau, p.use_groups, p.use_pattern, g, st); break;
    case 5: run_keeper(m, pat, p.tau, p.row_block_size, p.use_pattern, g, st); break;
  }
}

}  // namespace

const char* host_blocking(int64_t rows, int64_t cols, const int64_t* rowptr, const int64_t* colind,
                          const BlockingParams& p, int64_t* grouping, BlockingStats* stats) {
  if (rows < 0 || cols <= 0) return "invalid matrix shape";
  if (p.block_col_size <= 0) return "column block size must be positive";
  if (p.sim_measure < 0 || p.sim_measure > 3) return "similarity measure must be 0..3";
  if ((p.algo == 2 || p.algo == 5 || p.force_fixed_size) && p.row_block_size <= 0)
    return "row block size must be positive";
  Acc st;
  if (p.algo == 2) {  // FixedBlocking (blocking.cpp:554-562)
    for (int64_t i = 0; i < rows; ++i) grouping[i] = i / p.row_block_size;
  } else if (p.algo == 0 || p.algo == 3 || p.algo == 4 || p.algo == 5) {
    Rows m;
    m.rows = rows; m.cols = cols; m.w = p.block_col_size;
    m.block_cols = (cols - 1) / m.w + 1;
    m.rowptr = rowptr; m.colind = colind;
    for (int64_t i = 0; i < rows; ++i) {
      if (rowptr[i + 1] < rowptr[i]) return "rowptr must be non-decreasing";
      for (int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q) {
        if (colind[q] < 0 || colind[q] >= cols) return "column index out of range";
        if (q > rowptr[i] && colind[q] <= colind[q - 1]) return "columns must be strictly ascending inside a row";
      }
    }
    std::fill(grouping, grouping + rows, static_cast<int64_t>(-1));
    if (p.sim_measure == 1 && !p.force_list_model) {
      if (m.block_cols > INT32_MAX) return "too many column blocks";
      m.bptr.assign(rows + 1, 0);
      m.bidx.reserve(static_cast<size_t>(rowptr[rows]));
      for (int64_t i = 0; i < rows; ++i) {
        int64_t last = -1;
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q) {
          const int64_t b = colind[q] / m.w;
          if (b != last) { m.bidx.push_back(static_cast<int32_t>(b)); last = b; }
        }
        m.bptr[i + 1] = static_cast<int64_t>(m.bidx.size());
      }
      BlockPattern pat(m);
      dispatch(m, pat, p, grouping, st);
    } else {
      ListPattern pat(m, p.sim_measure);
      dispatch(m, pat, p, grouping, st);
    }
  } else {
    return "blocking algorithm not supported (0 iterative, 2 fixed, 3 clocked, 4 queue, 5 max-size)";
  }
  if (p.force_fixed_size && p.algo != 2) {  // get_fixed_size_grouping (utilities.cpp:45-54)
    std::vector<int64_t> perm(rows);
    host_permutation(grouping, rows, perm.data());
    for (int64_t i = 0; i < rows; ++i) grouping[perm[i]] = i / p.row_block_size;
  }
  if (stats) {
    stats->comparisons = st.comparisons;
    stats->merges = st.merges;
    if (p.algo == 3 || p.algo == 4 || p.algo == 5) {   // 0/0 = NaN like the reference (:239-240)
      stats->average_merge_tau = st.merge_tau / st.merges;
      stats->average_row_distance = st.row_distance / st.merges;
    } else {
      stats->average_merge_tau = stats->average_row_distance = 0;
    }
  }
  return "";
}

}  // namespace sparta
