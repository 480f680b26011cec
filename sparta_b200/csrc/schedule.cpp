// Host tile scheduler (see schedule.h).
//
// What it replaces in the reference: the per-call cursor arrays
// jab_positions / mab_positions and the "nzs-th block of every block-row"
// double loop of cublas_fixed_blocks_multiply (src/cuda/cuda_utilities.cpp:108-182)
// and the per-level pointer arrays of cublas_blockmat_batched (:811-856).
#include "host_formats.h"
#include "schedule.h"

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <queue>
#include <thread>
#include <utility>

namespace sparta {

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

static const char* build_structure_for(const BlockRows& br, const ScheduleOptions& opt, int64_t chain,
                                       Structure* out);

// Cuts chunks [c0, c1) of the super-row starting at chunks[begin] into passes such that no
// ACCUMULATOR receives more than `chain` MMAs in one pass (balanced; chain <= 0: one pass).  A
// member's accumulator only takes the MMAs of the chunks it is present in, so the count is kept
// per member: a super-row of many short, rarely overlapping block-rows needs far fewer passes than
// its chunk count suggests.  Appends the first chunk of every pass to *offs.
static void cut_passes(const std::vector<Chunk>& chunks, int64_t begin, int32_t c0, int32_t c1,
                       int64_t chain, std::vector<int32_t>* offs, int64_t* longest) {
  int64_t total[kMaxMembers] = {0};
  for (int32_t c = c0; c < c1; ++c) {
    const Chunk& ch = chunks[begin + c];
    for (uint32_t m = ch.mask; m; m &= m - 1) total[__builtin_ctz(m)] += ch.ksteps;
  }
  int64_t worst = 0;
  for (int m = 0; m < kMaxMembers; ++m) worst = std::max(worst, total[m]);
  const int64_t passes = chain > 0 ? std::max<int64_t>(1, (worst + chain - 1) / chain) : 1;
  const int64_t target = (worst + passes - 1) / passes;
  int64_t run[kMaxMembers] = {0};
  int64_t run_max = 0;
  offs->push_back(c0);
  for (int32_t c = c0; c < c1; ++c) {
    const Chunk& ch = chunks[begin + c];
    int64_t next_max = run_max;
    for (uint32_t m = ch.mask; m; m &= m - 1) next_max = std::max(next_max, run[__builtin_ctz(m)] + ch.ksteps);
    if (run_max > 0 && next_max > target) {
      if (longest) *longest = std::max(*longest, run_max);
      offs->push_back(c);
      for (int m = 0; m < kMaxMembers; ++m) run[m] = 0;
      run_max = 0;
    }
    for (uint32_t m = ch.mask; m; m &= m - 1) {
      int64_t& r = run[__builtin_ctz(m)];
      r += ch.ksteps;
      run_max = std::max(run_max, r);
    }
  }
  if (longest) *longest = std::max(*longest, run_max);
}

// chain limit in MMAs per accumulator: tf32 sums are held to <= 1e-5, and the tensor
// core's fp32 accumulation truncates (measured: 5.6e-5 relative on an all-positive chain of ~8000
// MMAs), so tf32 bounds the chain by default; bf16 / fp16 (tolerance 2e-2) do not.
const char* build_structure(const BlockRows& br, const ScheduleOptions& opt, Structure* out) {
  int64_t chain = opt.max_chain;
  if (chain == 0) chain = opt.precision == PREC_TF32 ? 256 : -1;
  if (chain > 0 && chain < 8) return "max_chain must be at least 8";
  // First try the requested accumulator width with unbounded chains; only if some super-row
  // exceeds the limit rebuild with 256-column super-rows (the other 256 TMEM columns then hold
  // the master copy) cut into passes.
  const char* e = build_structure_for(br, opt, -1, out);
  if (*e || chain < 0 || out->max_chain_seen <= chain) return e;
  ScheduleOptions o2 = opt;
  o2.acc_cols = 256;
  o2.tiles = 1;            // bounded chains keep the master accumulators in the other half of TMEM
  return build_structure_for(br, o2, chain, out);
}

static const char* build_structure_for(const BlockRows& br, const ScheduleOptions& opt, int64_t chain,
                                       Structure* out) {
  const bool timing = getenv("SPARTA_TIMING") != nullptr;
  const auto tp0 = std::chrono::steady_clock::now();
  auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tp0).count(); };
  Structure& st = *out;
  st = Structure();
  st.pair = opt.pair ? 1 : 0;
  st.tiles = opt.tiles;
  st.acc_cols = opt.acc_cols;
  st.master_col = chain > 0 ? 256 : 0;
  st.chain = chain;
  st.sparse_images = !br.sub_ptr.empty();
  if (br.w <= 0) return "column block size must be positive";
  if (opt.seg_rows < 16 || opt.seg_rows > 256 || opt.seg_rows % 16) return "seg_rows must be a multiple of 16 in [16,256]";
  if (opt.tiles != 1 && opt.tiles != 2 && opt.tiles != 4) return "tiles must be 1, 2 or 4";
  if (opt.tiles > 1 && (opt.acc_cols != 512 / opt.tiles || chain > 0))
    return "wide items need acc_cols = 512 / tiles and unbounded accumulation chains";
  if (opt.acc_cols != 128 && opt.acc_cols != 256 && opt.acc_cols != 512) return "acc_cols must be 128, 256 or 512";
  if (opt.seg_rows > opt.acc_cols) return "seg_rows exceeds acc_cols";
  const int esize = prec_esize(opt.precision);
  const int katom = 128 / esize;    // k elements per 128-byte swizzle row
  const int kstep = 32 / esize;     // k elements per MMA (K = 16 for 16-bit, 8 for tf32)
  const int kalign = 16 / esize;    // TMA needs a 16-byte aligned start along k
  const int nshare = st.pair ? 2 : 1;
  // tcgen05 instruction descriptor (kind::f16 / kind::tf32): D fp32 at [4,6), A/B format at
  // [7,10)/[10,13) (0 f16, 1 bf16, 2 tf32), both operands K-major, N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t fmt = opt.precision == PREC_BF16 ? 1u : (opt.precision == PREC_FP16 ? 0u : 2u);
  const uint32_t idesc_base = (1u << 4) | (fmt << 7) | (fmt << 10) | (((st.pair ? 256u : 128u) >> 4) << 24);

  // 0. order of the block-rows.  C rows are written wherever row_part says, so the order in
  //    which block-rows are grouped into super-rows is free: putting block-rows with similar
  //    nonzero-block counts together makes their column-block lists overlap (on R-MAT the
  //    dense rows are near-supersets of each other), so a B panel feeds more MMA rows.
  const int64_t nb = br.count();
  std::vector<int64_t> order(nb);
  for (int64_t b = 0; b < nb; ++b) order[b] = b;
  if (opt.sort_rows)
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
      return br.ptr[a + 1] - br.ptr[a] > br.ptr[b + 1] - br.ptr[b];
    });

  // 1. segments, tagged with their block-row and row offset inside it
  struct SegSrc { int64_t b; int64_t row_off; };
  std::vector<SegSrc> seg_src;
  for (int64_t oi = 0; oi < nb; ++oi) {
    const int64_t b = order[oi];
    const int64_t H = br.height[b];
    const int64_t nblk = br.ptr[b + 1] - br.ptr[b];
    st.n_blocks += nblk;  // zero-height block-rows keep their (empty) blocks in the count
    if (H <= 0) continue;
    st.rows = std::max<int64_t>(st.rows, br.row0[b] + H);
    if (br.row0[b] + H > INT32_MAX) return "shard has more than 2^31 rows";
    if (!br.sub_ptr.empty()) {
      // fused block-rows: blocks and area of the ORIGINAL blocks
      st.n_blocks -= nblk;
      for (int64_t q = br.ptr[b]; q < br.ptr[b + 1]; ++q) {
        st.n_blocks += br.sub_ptr[q + 1] - br.sub_ptr[q];
        for (int64_t t = br.sub_ptr[q]; t < br.sub_ptr[q + 1]; ++t) st.nztot += br.sub_h[t] * br.w;
      }
    } else if (br.blk_kw.empty()) st.nztot += nblk * H * br.w;
    else for (int64_t q = br.ptr[b]; q < br.ptr[b + 1]; ++q) st.nztot += H * br.blk_kw[q];
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

  // 2. super-rows: consecutive segments packed into one accumulator stage (sequential: cheap, and it
  //    assigns the accumulator columns of every segment)
  const size_t nseg = st.segs.size();
  struct Group { size_t s0, s1; int cols; uint32_t break_mask; int64_t blocks; };
  std::vector<Group> groups;
  for (size_t s0 = 0; s0 < nseg;) {
    int cols = 0;
    size_t s1 = s0;
    int64_t blocks = 0;
    while (s1 < nseg && (s1 - s0) < static_cast<size_t>(kMaxMembers) &&
           cols + st.segs[s1].h_pad <= opt.acc_cols) {
      st.segs[s1].tmem_col = cols;
      cols += st.segs[s1].h_pad;
      blocks += br.ptr[seg_src[s1].b + 1] - br.ptr[seg_src[s1].b];
      ++s1;
    }
    // fixed cuts so that no MMA run spans more than 256 accumulator columns
    uint32_t break_mask = 0;
    for (size_t s = s0, run_cols = 0; s < s1; ++s) {
      if (run_cols + st.segs[s].h_pad > 256) {
        break_mask |= 1u << (s - s0);
        run_cols = 0;
      }
      run_cols += st.segs[s].h_pad;
    }
    groups.push_back({s0, s1, cols, break_mask, blocks});
    s0 = s1;
  }

  // 3.-4. per super-row: merged column-block list, chunks, run tables, pack jobs.  Super-rows are
  //    independent, so ranges of them (balanced on block count) are built by all host threads into
  //    partial structures whose offsets are rebased when they are appended in order: the scheduler sits
  //    on the critical path of every create (31 ms single-threaded at BASELINE config #3).
  const std::vector<Segment>& segs = st.segs;
  auto build_range = [&](size_t g_lo, size_t g_hi, Structure& part) -> const char* {
  Structure& st = part;     // everything below appends to the partial structure
  std::vector<std::pair<int64_t, int>> merged;  // (jb, member)
  int32_t cols_of[kMaxMembers + 1];
  const int64_t* slot_of[kMaxMembers];          // per member: position of jb inside its block-row
  for (size_t gi = g_lo; gi < g_hi; ++gi) {
    const size_t s0 = groups[gi].s0, s1 = groups[gi].s1;
    const int cols = groups[gi].cols;
    const uint32_t break_mask = groups[gi].break_mask;
    SuperRow sr{};
    sr.seg_begin = static_cast<int32_t>(s0);
    for (size_t s = s0; s < s1; ++s) cols_of[s - s0] = segs[s].tmem_col;
    cols_of[s1 - s0] = cols;
    sr.break_mask = break_mask;
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
    // Per work item: draining the accumulator (measured 20.4 k cycles for 512 columns while the
    // other SMs keep the L2 busy; the stores, not the TMEM reads, are the limit -- see
    // scripts/microbench/epilogue_rate.cu) plus the tensor pipe running dry and refilling.
    const double fixed = 16000.0 + 40.0 * cols * opt.tiles;   // a wide item drains `tiles` copies of the columns
    double cost = fixed;
    size_t i = 0;
    while (i < merged.size()) {
      const int64_t jb = merged[i].first;
      size_t i1 = i;
      uint32_t mask = 0;
      while (i1 < merged.size() && merged[i1].first == jb) {
        const int m = merged[i1].second;
        mask |= 1u << m;
        const int64_t b = seg_src[s0 + m].b;
        const int64_t* cb = br.col.data() + br.ptr[b];
        slot_of[m] = cb + (std::lower_bound(cb, br.col.data() + br.ptr[b + 1], jb) - cb);
        ++i1;
      }
      uint32_t rows_present = 0;
      for (size_t t = i; t < i1; ++t) rows_present += segs[s0 + merged[t].second].h_pad;
      // K slabs of this column block: start at the 16-byte aligned k at or below jb*w
      const int64_t q_any = slot_of[merged[i].second] - br.col.data();   // same k range for every member
      const int64_t kblk = br.blk_k0.empty() ? jb * br.w : br.blk_k0[q_any];
      const int64_t kw = br.blk_kw.empty() ? br.w : br.blk_kw[q_any];
      const int64_t ka = kblk / kalign * kalign;
      const int shift = static_cast<int>(kblk - ka);
      const int64_t atoms = (shift + kw + katom - 1) / katom;
      for (int64_t a = 0; a < atoms; ++a) {
        const int64_t k_lo = a * katom - shift;  // block-local k of image column 0
        const int k_used = static_cast<int>(std::min<int64_t>(katom, kw - k_lo));
        Chunk ch{};
        const int64_t k0 = ka + a * katom;
        if (k0 > INT32_MAX) return "k index exceeds 2^31";
        ch.k0 = static_cast<int32_t>(k0);
        ch.mask = mask;
        ch.ksteps = (k_used + kstep - 1) / kstep;
        if ((st.a_bytes >> 4) > UINT32_MAX) return "packed A exceeds 64 GiB";
        ch.a_off16 = static_cast<uint32_t>(st.a_bytes >> 4);
        const uint32_t bytes = rows_present * 128u;
        const uint32_t share = bytes / nshare;       // bytes each CTA stages for this chunk
        uint32_t cursor[2] = {0, 0};                 // write position inside each CTA's share
        const uint64_t chunk_base = st.a_bytes;
        const size_t tbl = st.tables.size();
        st.tables.resize(tbl + kTableWords, 0u);
        uint32_t nruns = 0, rows_before = 0;   // rows_before: per-CTA rows of the earlier runs
        // Images are laid out run by run (the unit of one MMA); in pair mode the first half of a
        // run's rows goes to CTA 0's share and the second half to CTA 1's (cta_group::2 reads
        // N/2 rows of the N operand from each CTA).
        for_each_run(mask, break_mask, cols_of, [&](int mb, int me, int N) {
          const int half = N / nshare;
          st.tables[tbl + 4 + 2 * nruns] = idesc_base | (static_cast<uint32_t>(N >> 3) << 17);
          st.tables[tbl + 5 + 2 * nruns] = (static_cast<uint32_t>(cols_of[mb]) << 16) | ((rows_before * 128u) >> 4);
          ++nruns;
          rows_before += static_cast<uint32_t>(half);
          for (int m = mb; m < me; ++m) {
            const size_t s = s0 + m;
            const int64_t b = seg_src[s].b;
            const int64_t q = slot_of[m] - br.col.data();
            const int lo = cols_of[m] - cols_of[mb];            // member rows inside the run
            const int hi = lo + segs[s].h_pad;
            for (int cta = 0; cta < nshare; ++cta) {
              const int r0 = std::max(lo, cta * half), r1 = std::min(hi, (cta + 1) * half);
              if (r0 >= r1) continue;
              if (!br.sub_ptr.empty()) {
                // one job per sub-block that reaches into rows [r0, r1) of the run; a job writes
                // only its own rows (the buffer is zeroed first, Structure::sparse_images)
                const int64_t b_lo = seg_src[s].row_off + (r0 - lo), b_hi = seg_src[s].row_off + (r1 - lo);
                for (int64_t t = br.sub_ptr[q]; t < br.sub_ptr[q + 1]; ++t) {
                  const int64_t a = std::max(b_lo, br.sub_off[t]);
                  const int64_t e = std::min(b_hi, br.sub_off[t] + br.sub_h[t]);
                  if (a >= e) continue;
                  PackJob job;
                  job.src_rs = br.sub_rs[t];
                  job.src_ks = br.sub_ks[t];
                  job.src_base = br.sub_src[t] + (a - br.sub_off[t]) * job.src_rs;
                  job.h = job.h_pad = static_cast<int32_t>(e - a);
                  job.k_lo = static_cast<int32_t>(k_lo);
                  job.k_w = static_cast<int32_t>(kw);
                  job.r_base = static_cast<int32_t>(a - b_lo);
                  job.pad_[0] = job.pad_[1] = 0;
                  job.dst_off16 = static_cast<uint32_t>((chunk_base + cta * share + cursor[cta]) >> 4);
                  st.jobs.push_back(job);
                }
                cursor[cta] += static_cast<uint32_t>(r1 - r0) * 128u;
                continue;
              }
              PackJob job;
              job.src_rs = br.blk_rs.empty() ? br.rs[b] : br.blk_rs[q];
              job.src_ks = br.blk_ks.empty() ? br.ks[b] : br.blk_ks[q];
              job.src_base = br.src[q] + (seg_src[s].row_off + (r0 - lo)) * job.src_rs;
              job.h = std::max(0, std::min(segs[s].h - (r0 - lo), r1 - r0));
              job.h_pad = r1 - r0;
              job.k_lo = static_cast<int32_t>(k_lo);
              job.k_w = static_cast<int32_t>(kw);
              job.r_base = 0;
              job.pad_[0] = job.pad_[1] = 0;
              job.dst_off16 = static_cast<uint32_t>((chunk_base + cta * share + cursor[cta]) >> 4);
              st.jobs.push_back(job);
              cursor[cta] += static_cast<uint32_t>(r1 - r0) * 128u;
            }
          }
        });
        st.tables[tbl] = nruns;
        st.tables[tbl + 1] = static_cast<uint32_t>(ch.ksteps);
        ch.tbl_bytes = 16u + 8u * ((nruns + 1u) & ~1u);   // multiple of 16 for the bulk copy
        if ((tbl >> 2) > UINT32_MAX) return "run tables exceed 64 GiB";
        ch.tbl_off16 = static_cast<uint32_t>(tbl >> 2);
        st.tables.resize(tbl + ch.tbl_bytes / 4);
        ch.a_bytes = bytes;
        st.a_bytes += bytes;
        st.max_chunk_bytes = std::max(st.max_chunk_bytes, share);
        st.chunks.push_back(ch);
        // modelled cycles per CTA: tensor pipe N/2 per K step; L2 -> smem: fitted to the kernel
        // times of 30 shards of the bench matrix (scripts/fit_cost_model.py): 511 cycles per chunk
        // + 1.25 per staged row of A in pair mode (64 bytes per row and CTA => ~50 B/cycle/SM),
        // 36.7 k per item
        const double tensor = ch.ksteps * (rows_present * 0.5) * opt.tiles;
        const double memory = 183.0 + (opt.tiles * kPanelBytes + share) / 50.0;
        st.chunk_cost.push_back(static_cast<float>(std::max(tensor, memory)));
        cost += st.chunk_cost.back();
      }
      i = i1;
    }
    sr.chunk_count = static_cast<int32_t>(st.chunks.size()) - sr.chunk_begin;
    // passes: balanced cuts so that no pass issues more than `chain` MMAs into an accumulator
    st.pass_ptr.push_back(static_cast<int32_t>(st.pass_off.size()));
    cut_passes(st.chunks, sr.chunk_begin, 0, sr.chunk_count, chain, &st.pass_off, &st.max_chain_seen);
    st.srows.push_back(sr);
    st.srow_cost.push_back(cost);
    st.srow_fixed.push_back(fixed);
  }
  return "";
  };   // build_range

  const double t_groups = since();
  // ranges of super-rows balanced on block count, one per host thread
  const size_t n_groups = groups.size();
  int T = host_thread_budget(16);
  if (st.n_blocks < 20000 || n_groups < 8) T = 1;
  T = static_cast<int>(std::min<size_t>(T, std::max<size_t>(n_groups, 1)));
  std::vector<size_t> cut(T + 1, 0);
  {
    int64_t total = 0, run = 0;
    for (const Group& g : groups) total += g.blocks + 8;
    int t = 1;
    for (size_t gi = 0; gi < n_groups && t < T; ++gi) {
      run += groups[gi].blocks + 8;
      if (run * T >= total * t) cut[t++] = gi + 1;
    }
    for (; t < T; ++t) cut[t] = n_groups;
    cut[T] = n_groups;
  }
  std::vector<Structure> parts(T);
  std::vector<const char*> errs(T, "");
  if (T == 1) {
    errs[0] = build_range(0, n_groups, parts[0]);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t) pool.emplace_back([&, t] { errs[t] = build_range(cut[t], cut[t + 1], parts[t]); });
    for (auto& th : pool) th.join();
  }
  for (int t = 0; t < T; ++t)
    if (*errs[t]) return errs[t];
  const double t_built = since();
  // Append the partial structures in order, rebasing their offsets.  The pack jobs (64 bytes per image:
  // 22 MB at BASELINE config #3) are rebased in place by all threads and kept as parts -- copying them
  // into one array cost more than building them.
  std::vector<uint64_t> base_a(T + 1, 0);
  std::vector<size_t> base_tbl(T + 1, 0), base_chunk(T + 1, 0), base_pass(T + 1, 0);
  for (int t = 0; t < T; ++t) {
    base_a[t + 1] = base_a[t] + parts[t].a_bytes;
    base_tbl[t + 1] = base_tbl[t] + parts[t].tables.size();
    base_chunk[t + 1] = base_chunk[t] + parts[t].chunks.size();
    base_pass[t + 1] = base_pass[t] + parts[t].pass_off.size();
  }
  if ((base_a[T] >> 4) > UINT32_MAX) return "packed A exceeds 64 GiB";
  if ((base_tbl[T] >> 2) > UINT32_MAX) return "run tables exceed 64 GiB";
  if (base_chunk[T] > static_cast<size_t>(INT32_MAX)) return "too many chunks";
  {
    auto rebase = [&](int t) {
      Structure& pt = parts[t];
      const uint32_t da = static_cast<uint32_t>(base_a[t] >> 4), dt = static_cast<uint32_t>(base_tbl[t] >> 2);
      for (Chunk& ch : pt.chunks) { ch.a_off16 += da; ch.tbl_off16 += dt; }
      for (PackJob& job : pt.jobs) job.dst_off16 += da;
      for (SuperRow& sr : pt.srows) sr.chunk_begin += static_cast<int32_t>(base_chunk[t]);
      for (int32_t& pp : pt.pass_ptr) pp += static_cast<int32_t>(base_pass[t]);
    };
    if (T == 1) {
      rebase(0);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < T; ++t) pool.emplace_back(rebase, t);
      for (auto& th : pool) th.join();
    }
  }
  st.chunks.reserve(base_chunk[T]);
  st.tables.reserve(base_tbl[T]);
  st.chunk_cost.reserve(base_chunk[T]);
  st.pass_off.reserve(base_pass[T]);
  st.job_parts.clear();
  for (int t = 0; t < T; ++t) {
    Structure& pt = parts[t];
    st.chunks.insert(st.chunks.end(), pt.chunks.begin(), pt.chunks.end());
    st.tables.insert(st.tables.end(), pt.tables.begin(), pt.tables.end());
    st.chunk_cost.insert(st.chunk_cost.end(), pt.chunk_cost.begin(), pt.chunk_cost.end());
    st.pass_off.insert(st.pass_off.end(), pt.pass_off.begin(), pt.pass_off.end());
    st.pass_ptr.insert(st.pass_ptr.end(), pt.pass_ptr.begin(), pt.pass_ptr.end());
    st.srows.insert(st.srows.end(), pt.srows.begin(), pt.srows.end());
    st.srow_cost.insert(st.srow_cost.end(), pt.srow_cost.begin(), pt.srow_cost.end());
    st.srow_fixed.insert(st.srow_fixed.end(), pt.srow_fixed.begin(), pt.srow_fixed.end());
    st.n_jobs += static_cast<int64_t>(pt.jobs.size());
    st.job_parts.emplace_back(std::move(pt.jobs));
    st.a_bytes += pt.a_bytes;
    st.max_chunk_bytes = std::max(st.max_chunk_bytes, pt.max_chunk_bytes);
    st.max_chain_seen = std::max(st.max_chain_seen, pt.max_chain_seen);
  }
  st.pass_ptr.push_back(static_cast<int32_t>(st.pass_off.size()));
  if (st.chunks.size() > static_cast<size_t>(INT32_MAX)) return "too many chunks";
  if (timing)
    fprintf(stderr, "sparta schedule: order + segments + groups %.1f ms, %d threads build %.1f, merge %.1f (%zu chunks, %zu pack jobs)\n",
            t_groups, T, t_built - t_groups, since() - t_built, st.chunks.size(), static_cast<size_t>(st.n_jobs));
  return "";
}

// Work order and static assignment.
//
// Columns of B are processed in GROUPS of `group_tiles` column tiles whose B slab
// (cols x group width) stays L2-resident while every super-row streams its A images past it,
// so HBM sees B about once and A once per group instead of once per column tile.  Inside a
// group the workers form TEAMS of `team` = group_tiles workers: a team walks the same sequence
// of super-rows side by side, one column tile each, so a super-row's A images are fetched from
// HBM by whichever member gets there first and hit in L2 for the others.
//
// Two ways of handing the (group, super-row) units to the teams:
//   whole units -- list-scheduled, heaviest first, onto the team that frees up first.  Right when
//     there are many more units than teams (a whole matrix on one GPU: 516 units for 37 teams).
//   split units -- the units' chunk lists are laid end to end and cut into one equal-cost piece
//     per team ("stream-K" along the block-row's column-block list).  Right when a shard holds few
//     super-rows (one rank of an 8-GPU run: ~40 units for 37 teams, max/mean load 1.6-1.9 with
//     whole units).  A piece of a cut unit adds its partial sums to C with fp32 reductions.
namespace {

struct TeamPiece { int32_t group, srow, c0, c1; bool partial; };

struct TeamPlan {
  int workers = 0, team = 1;
  int spare = 0;   // width of one extra, narrower team made of the workers a multiple of `team` leaves over
  std::vector<std::vector<TeamPiece>> per_team;
  double max_cost = 0, mean_cost = 0;
  int cut_units = 0;
};

// Team width = column tiles per group: the largest divisor of the tile count whose B slab fits.
// The worker count is rounded DOWN to a multiple of the team width (74 CTA pairs run as 18 teams
// of 4): measured on the bench matrix, teams of 4 on 144 CTAs beat teams of 2 on 148 by 4-5 % on
// quarter and eighth shards and tie on the whole matrix -- wider teams halve the number of groups,
// i.e. the work items per worker and the passes of A through HBM.
int pick_team(int* workers, int64_t tiles, int64_t fit) {
  int team = 1;
  for (int t = 1; t <= *workers && t <= fit && t <= tiles; ++t)
    if (tiles % t == 0) team = t;
  *workers = *workers / team * team;
  return team;
}

void finish_costs(TeamPlan* tp, const std::vector<double>& load) {
  double total = 0;
  for (double l : load) {
    tp->max_cost = std::max(tp->max_cost, l);
    total += l;
  }
  tp->mean_cost = load.empty() ? 0.0 : total / load.size();
}

// The workers a multiple of the team width leaves over (74 CTA pairs = 9 teams of 8 + 2) form one
// narrower team: its members take the column tiles of a group in turns (4 each for 2 members of
// an 8-tile group), so a unit costs it team / spare times as much and the list scheduler gives it
// proportionally fewer units.
void plan_whole(const Structure& st, const std::vector<int32_t>& by_cost, int64_t groups, TeamPlan* tp) {
  const int n_full = tp->workers / tp->team;
  const int n_teams = n_full + (tp->spare > 0 ? 1 : 0);
  const double spare_factor = tp->spare > 0 ? static_cast<double>(tp->team) / tp->spare : 1.0;
  tp->per_team.assign(n_teams, {});
  typedef std::pair<double, int> Load;
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int t = 0; t < n_teams; ++t) heap.push(Load(0.0, t));
  for (int64_t g = 0; g < groups; ++g)
    for (int32_t s : by_cost) {
      Load l = heap.top();
      heap.pop();
      const double unit = st.srow_cost[s] * (l.second >= n_full ? spare_factor : 1.0);
      if (l.second >= n_full) {
        // would the narrow team finish this unit later than a full team could?  then pass
        Load alt = heap.top();
        if (alt.first + st.srow_cost[s] < l.first + unit) {
          heap.pop();
          heap.push(l);
          l = alt;
          tp->per_team[l.second].push_back({static_cast<int32_t>(g), s, 0, st.srows[s].chunk_count, false});
          l.first += st.srow_cost[s];
          heap.push(l);
          continue;
        }
      }
      tp->per_team[l.second].push_back({static_cast<int32_t>(g), s, 0, st.srows[s].chunk_count, false});
      l.first += unit;
      heap.push(l);
    }
  std::vector<double> load;
  while (!heap.empty()) { load.push_back(heap.top().first); heap.pop(); }
  finish_costs(tp, load);
}

// Greedy fill of ONE column group: teams take the super-rows in sequence until they hold `target`
// cycles; a super-row that does not fit is cut at the chunk where the team is full.  The last
// team takes whatever is left.
void fill_split(const Structure& st, const std::vector<int32_t>& by_cost, const std::vector<double>& prefix,
                const std::vector<double>& slow, double target, std::vector<std::vector<TeamPiece>>* per_team,
                std::vector<double>* load) {
  // slow[t]: how many times longer team t takes for the same piece (1 for a full team, team / spare
  // for the narrow team of leftover workers); loads and the target are times
  const int n_teams = static_cast<int>(slow.size());
  const int32_t kMinChunks = 8;        // never cut off a piece shorter than this
  per_team->assign(n_teams, {});
  load->assign(n_teams, 0.0);
  int t = 0;
  for (int32_t s : by_cost) {
    const SuperRow& sr = st.srows[s];
    const double* pf = prefix.data() + sr.chunk_begin;   // pf[c] - pf[0] = cost of the chunks before c
    const double fixed = st.srow_fixed[s];
    int32_t c = 0;
    if (sr.chunk_count == 0) {   // no blocks at all: the item only writes zeros
      if (t < n_teams - 1 && (*load)[t] + fixed * slow[t] > target) ++t;
      (*per_team)[t].push_back({0, s, 0, 0, false});
      (*load)[t] += fixed * slow[t];
    }
    while (c < sr.chunk_count) {
      const double rest = pf[sr.chunk_count] - pf[c];
      const double room = (target - (*load)[t]) / slow[t] - fixed;
      if (t == n_teams - 1 || rest <= room) {
        (*per_team)[t].push_back({0, s, c, sr.chunk_count, false});
        (*load)[t] += (fixed + rest) * slow[t];
        break;
      }
      // largest e with cost of [c, e) <= room
      const int32_t e = static_cast<int32_t>(
          std::upper_bound(pf + c, pf + sr.chunk_count + 1, pf[c] + std::max(room, 0.0)) - pf) - 1;
      if (e - c >= kMinChunks && sr.chunk_count - e >= kMinChunks) {
        (*per_team)[t].push_back({0, s, c, e, true});
        (*load)[t] += (fixed + (pf[e] - pf[c])) * slow[t];
        c = e;
      }
      ++t;   // this team is full (or cannot take a piece worth cutting)
    }
  }
  // a piece is partial iff it is not the whole chunk list of its super-row
  for (auto& list : *per_team)
    for (auto& pc : list) pc.partial = pc.c0 != 0 || pc.c1 != st.srows[pc.srow].chunk_count;
}

// The cut is made inside ONE column group and repeated for every group, so that all teams stay in
// the same group at the same time and its B slab stays L2-resident (cutting the concatenation of
// all groups instead put the teams in different groups: measured 30 % slower at 2 shards, the B
// panels came from HBM).  Odd groups hand the pieces out in reverse team order to even out what
// the greedy fill leaves to the last team.  The workers a multiple of the team width leaves over
// form one narrower team (always the last one) that gets proportionally shorter pieces.
void plan_split(const Structure& st, const std::vector<int32_t>& by_cost, int64_t groups, TeamPlan* tp) {
  const int n_full = tp->workers / tp->team;
  const int n_teams = n_full + (tp->spare > 0 ? 1 : 0);
  const double factor = tp->spare > 0 ? static_cast<double>(tp->team) / tp->spare : 1.0;
  std::vector<double> slow_fwd(n_teams, 1.0), slow_rev(n_teams, 1.0);
  if (tp->spare > 0) { slow_fwd[n_teams - 1] = factor; slow_rev[0] = factor; }
  std::vector<double> prefix(st.chunk_cost.size() + 1, 0.0);
  for (size_t c = 0; c < st.chunk_cost.size(); ++c) prefix[c + 1] = prefix[c] + st.chunk_cost[c];
  double total = 0, biggest = 0;
  for (int32_t s : by_cost) { total += st.srow_cost[s]; biggest = std::max(biggest, st.srow_cost[s]); }
  // Every cut re-pays the fixed part of the unit it cuts and pieces have a minimum length, so the
  // load the greedy fill ends up with is not a monotone function of the target: scan a ladder of
  // targets from the ideal mean upwards and keep the fill with the smallest maximum load.
  const double lo = total / (n_full + (tp->spare > 0 ? 1.0 / factor : 0.0));
  const double hi = std::max(2.0 * lo, lo + biggest * factor);
  std::vector<std::vector<TeamPiece>> best[2], cur;
  std::vector<double> best_load[2], cur_load;
  double best_worst[2] = {0, 0};
  for (int dir = 0; dir < 2; ++dir) {
    const std::vector<double>& slow = dir ? slow_rev : slow_fwd;
    for (double target = lo; ; target *= 1.01) {
      const bool last = target >= hi;
      fill_split(st, by_cost, prefix, slow, last ? hi : target, &cur, &cur_load);
      const double worst = *std::max_element(cur_load.begin(), cur_load.end());
      if (best[dir].empty() || worst < best_worst[dir]) {
        best[dir].swap(cur);
        best_load[dir].swap(cur_load);
        best_worst[dir] = worst;
      }
      if (last) break;
    }
    if (tp->spare == 0) {   // symmetric: the reverse fill is the same lists
      best[1] = best[0];
      best_load[1] = best_load[0];
      break;
    }
  }
  tp->per_team.assign(n_teams, {});
  std::vector<double> load(n_teams, 0.0);
  for (int64_t g = 0; g < groups; ++g)
    for (int t = 0; t < n_teams; ++t) {
      const int dir = static_cast<int>(g & 1);
      const int src = dir ? n_teams - 1 - t : t;
      for (TeamPiece pc : best[dir][src]) {
        pc.group = static_cast<int32_t>(g);
        tp->per_team[t].push_back(pc);
        if (pc.partial && pc.c0 == 0) ++tp->cut_units;
      }
      load[t] += best_load[dir][src];
    }
  finish_costs(tp, load);
}

}  // namespace

const char* build_assignment(const Structure& st, const ScheduleOptions& opt, int64_t n,
                             int64_t k_total, Assignment* out) {
  Assignment& as = *out;
  as = Assignment();
  const int tile = (st.pair ? 2 * kTileJ : kTileJ) * st.tiles;   // columns of B one work item covers
  if (n <= 0 || n > INT32_MAX - tile) return "invalid number of B columns";
  const int64_t tiles = (n + tile - 1) / tile;
  const int64_t n_srows = static_cast<int64_t>(st.srows.size());
  if (static_cast<int64_t>(st.pass_off.size()) * tiles > INT32_MAX / 2) return "too many work items";
  const int all_workers = std::max(1, opt.num_ctas / (st.pair ? 2 : 1));
  as.cta_ptr.assign(1, 0);
  if (n_srows == 0) return "";

  // column tiles per group: as many as keep the group's slab of B (k_total x group width) within
  // the L2 budget.  The budget (default 160 MB) is deliberately above the 126 MB of L2: the workers
  // sweep k upwards roughly in step, so the panels in use at any time are a moving band of the
  // slab, not all of it.  Whole-unit plans, where every item starts at the head of a column-block
  // list, get twice the budget (measured on the bench matrix: the whole matrix is 1.3 % faster
  // with all of B, 268 MB, in one group; eighth shards under a split plan, whose pieces start
  // anywhere in the lists, are 4 % slower that way than with 134 MB groups).
  const double tile_bytes = static_cast<double>(k_total) * tile * prec_esize(opt.precision);
  const int64_t fit_whole = std::min<int64_t>(
      tiles, std::max<int64_t>(1, static_cast<int64_t>(2.0 * opt.l2_slab_bytes / std::max(tile_bytes, 1.0))));
  int64_t fit = std::max<int64_t>(1, static_cast<int64_t>(opt.l2_slab_bytes / std::max(tile_bytes, 1.0)));
  fit = std::min(fit, tiles);

  // Super-rows without a single nonzero block get no work item: their rows of C are zero after
  // the memset that follows every (re)allocation of C (set_B), nothing else ever writes them, and
  // with accumulate = 1 adding zero is a no-op.  (On the R-MAT bench matrix a fifth of the rows
  // are empty; draining an all-zero accumulator for each of them cost more than their share.)
  std::vector<int32_t> by_cost;
  for (int64_t s = 0; s < n_srows; ++s)
    if (st.srows[s].chunk_count > 0) by_cost.push_back(static_cast<int32_t>(s));
  if (by_cost.empty()) return "";
  std::stable_sort(by_cost.begin(), by_cost.end(),
                   [&](int32_t a, int32_t b) { return st.srow_cost[a] > st.srow_cost[b]; });

  TeamPlan whole, split;
  whole.workers = static_cast<int>(std::min<int64_t>(all_workers, static_cast<int64_t>(by_cost.size()) * tiles));
  {
    const int before = whole.workers;
    whole.team = pick_team(&whole.workers, tiles, fit_whole);
    // leftover workers: one narrower team whose width divides the team width
    for (int w2 = before - whole.workers; w2 >= 1; --w2)
      if (whole.team % w2 == 0) { whole.spare = w2; break; }
  }
  plan_whole(st, by_cost, (tiles + whole.team - 1) / whole.team, &whole);
  const TeamPlan* tp = &whole;
  if (opt.split != 1) {
    int64_t n_chunks = 0;
    for (const SuperRow& sr : st.srows) n_chunks += sr.chunk_count;
    // no more workers than pieces of >= 16 chunks
    split.workers = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(all_workers, n_chunks * tiles / 16)));
    {
      const int before = split.workers;
      split.team = pick_team(&split.workers, tiles, fit);
      for (int w2 = before - split.workers; w2 >= 1; --w2)
        if (split.team % w2 == 0) { split.spare = w2; break; }
    }
    plan_split(st, by_cost, (tiles + split.team - 1) / split.team, &split);
    const bool pays = split.max_cost < 0.95 * whole.max_cost;   // the worker that finishes last sets the time
    if (opt.split == 2 || pays) tp = &split;
  }

  as.workers = tp->workers + tp->spare;
  as.team = as.group_tiles = tp->team;
  as.grid = as.workers * (st.pair ? 2 : 1);
  as.max_cta_cost = tp->max_cost;
  as.mean_cta_cost = tp->mean_cost;
  std::vector<std::vector<int32_t>> per_worker(as.workers);
  std::vector<int32_t> offs;
  const size_t n_full_teams = static_cast<size_t>(tp->workers / tp->team);
  for (size_t t = 0; t < tp->per_team.size(); ++t)
    for (const TeamPiece& pc : tp->per_team[t]) {
      const int width = t < n_full_teams ? tp->team : tp->spare;      // members of this team
      const size_t team_base = (t < n_full_teams ? t : n_full_teams) * static_cast<size_t>(tp->team);
      const int64_t t0 = static_cast<int64_t>(pc.group) * tp->team;
      const int in_group = static_cast<int>(std::min<int64_t>(tp->team, tiles - t0));
      offs.clear();
      if (pc.partial) {
        cut_passes(st.chunks, st.srows[pc.srow].chunk_begin, pc.c0, pc.c1, st.chain, &offs, nullptr);
      } else {
        offs.assign(st.pass_off.begin() + st.pass_ptr[pc.srow], st.pass_off.begin() + st.pass_ptr[pc.srow + 1]);
      }
      for (int m = 0; m < in_group; ++m) {
        const int32_t j0 = static_cast<int32_t>((t0 + m) * tile);
        for (size_t ps = 0; ps < offs.size(); ++ps) {   // the passes of one piece stay together
          const int32_t off = offs[ps];
          const int32_t end = ps + 1 < offs.size() ? offs[ps + 1] : pc.c1;
          uint32_t count = static_cast<uint32_t>(end - off);
          if (ps > 0) count |= kItemNotFirst;
          if (ps + 1 < offs.size()) count |= kItemNotLast;
          else if (pc.partial) count |= kItemAtomic;
          per_worker[team_base + m % width].push_back(static_cast<int32_t>(as.items.size()));
          as.items.push_back(Item{pc.srow, j0, off, count});
        }
        if (pc.partial) {
          ++as.split_pieces;
          if (pc.c0 == 0) as.zero_jobs.push_back(ZeroJob{pc.srow, j0});
        }
      }
    }
  as.cta_ptr.assign(as.workers + 1, 0);
  as.cta_items.reserve(as.items.size());
  for (int c = 0; c < as.workers; ++c) {
    as.cta_ptr[c] = static_cast<int32_t>(as.cta_items.size());
    as.cta_items.insert(as.cta_items.end(), per_worker[c].begin(), per_worker[c].end());
  }
  as.cta_ptr[as.workers] = static_cast<int32_t>(as.cta_items.size());
  return "";
}

bool fuse_short_block_rows(const BlockRows& in, int max_rows, BlockRows* out) {
  if (!in.blk_k0.empty() || !in.sub_ptr.empty()) return false;
  const int64_t nb = in.count();
  // greedy runs of consecutive short block-rows
  std::vector<int64_t> first;   // first block-row of every output block-row, plus nb at the end
  bool any = false;
  for (int64_t b = 0; b < nb;) {
    first.push_back(b);
    int64_t e = b + 1;
    if (in.height[b] > 0 && in.height[b] < max_rows) {
      int64_t rows = in.height[b];
      while (e < nb && in.height[e] > 0 && rows + in.height[e] <= max_rows &&
             in.row0[e] == in.row0[e - 1] + in.height[e - 1]) {
        rows += in.height[e];
        ++e;
      }
    }
    if (e > b + 1) any = true;
    b = e;
  }
  first.push_back(nb);
  if (!any) return false;
  BlockRows f;
  f.w = in.w;
  f.ptr.push_back(0);
  f.sub_ptr.push_back(0);
  std::vector<std::pair<int64_t, int64_t>> ent;   // (column block, source block index)
  for (size_t g = 0; g + 1 < first.size(); ++g) {
    const int64_t b0 = first[g], b1 = first[g + 1];
    int64_t rows = 0;
    ent.clear();
    for (int64_t b = b0; b < b1; ++b) {
      for (int64_t q = in.ptr[b]; q < in.ptr[b + 1]; ++q) ent.emplace_back(in.col[q], q);
      rows += in.height[b];
    }
    f.row0.push_back(in.row0[b0]);
    f.height.push_back(rows);
    f.rs.push_back(in.rs[b0]);
    f.ks.push_back(in.ks[b0]);
    std::stable_sort(ent.begin(), ent.end());   // by column block, then by member (q ascending)
    // row offset of every member inside the fused block-row
    for (size_t i = 0; i < ent.size();) {
      size_t j = i;
      f.col.push_back(ent[i].first);
      f.src.push_back(in.src[ent[i].second]);
      while (j < ent.size() && ent[j].first == ent[i].first) {
        const int64_t q = ent[j].second;
        // the member that owns block q
        const int64_t b = std::upper_bound(in.ptr.begin() + b0, in.ptr.begin() + b1 + 1, q) - in.ptr.begin() - 1;
        f.sub_off.push_back(in.row0[b] - in.row0[b0]);
        f.sub_h.push_back(in.height[b]);
        f.sub_src.push_back(in.src[q]);
        f.sub_rs.push_back(in.rs[b]);
        f.sub_ks.push_back(in.ks[b]);
        ++j;
      }
      f.sub_ptr.push_back(static_cast<int64_t>(f.sub_off.size()));
      i = j;
    }
    f.ptr.push_back(static_cast<int64_t>(f.col.size()));
  }
  *out = std::move(f);
  return true;
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
