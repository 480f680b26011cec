"""Host tile scheduler + packer (csrc/schedule.cpp) checked on the CPU by interpreting the plan
with tests/sched_interp.py and comparing against the oracle's VBR::multiply restatement."""
import numpy as np
import pytest

import sparta_b200
from tests import sched_interp
from tests.util import random_vbr

CASES = [
    # block_rows, cols, w, heights, density, n, opts
    (6, 96, 16, [16] * 6, 0.5, 40, {}),
    (5, 70, 16, [3, 17, 1, 64, 30], 0.6, 130, {}),               # ragged heights, ragged cols, n tail
    (4, 64, 3, [4, 3, 1, 1], 0.7, 2, {}),                        # the TEST config's w = 3
    (3, 300, 100, [64, 64, 20], 0.8, 16, {}),                    # w > 64: two K slabs per block
    (7, 128, 64, [64] * 7, 0.4, 256, {"acc_cols": 256}),         # 4-wide super-rows, two stages
    (11, 256, 64, [64, 16, 64, 48, 64, 64, 32, 64, 64, 80, 64], 0.5, 300, {"row_order": 1}),  # runs split unevenly
    (2, 64, 32, [200, 70], 1.0, 8, {"seg_rows": 64}),            # tall block-rows split into segments
    (40, 64, 8, [1] * 40, 0.3, 8, {"acc_cols": 512}),            # many height-1 block-rows
    (3, 64, 16, [5, 0, 9], 0.9, 8, {}),                          # an empty block-row
    (7, 128, 64, [64] * 7, 0.4, 600, {"wide_tiles": 2}),         # wide items: 2 column tiles per item, 256 acc columns
    (11, 256, 64, [64, 16, 64, 48, 64, 64, 32, 64, 64, 80, 64], 0.5, 1100, {"wide_tiles": 4, "row_order": 1}),
    (40, 64, 8, [1] * 40, 0.3, 520, {"wide_tiles": 4}),          # 4 tiles, the last ones beyond n in pair mode
    (9, 512, 64, [64] * 9, 0.15, 1024, {}),                      # ER-like lists: the library picks wide items itself
]


@pytest.mark.parametrize("pair", [1, 2], ids=["single", "pair"])
@pytest.mark.parametrize("precision,esize", [("bf16", 2), ("tf32", 4)])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_plan_interpreter_matches_oracle(oracle, lib, case, precision, esize, pair):
    block_rows, cols, w, heights, density, n, opts = CASES[case]
    opts = dict(opts, cta_pair=pair)
    rng = np.random.default_rng(100 + case)
    v = random_vbr(rng, block_rows, cols, w, heights, density, values="int")
    Bm = rng.integers(-3, 4, size=(n, cols)).astype(np.float32)
    plan = sparta_b200.vbr_plan(v["rows"], cols, w, v["row_part"], v["nzcount"], v["jab"], n,
                                precision=precision, **opts)
    st = plan["stats"]
    assert st["nztot"] == v["mab"].size and st["nz_blocks"] == v["jab"].size
    assert st["rows"] == v["rows"]
    Cm = sched_interp.run_plan(plan, v["mab"], Bm, cols, n, v["rows"], esize=esize)
    Cref = oracle.vbr_multiply(v, Bm, n)
    assert not np.isnan(Cm).any(), "some C entries were never written"
    assert np.array_equal(Cm, Cref)  # small integers: exact in fp32 whatever the order


def test_plan_invariants(lib):
    rng = np.random.default_rng(7)
    heights = rng.integers(1, 130, size=60)
    v = random_vbr(rng, 60, 1000, 64, heights, 0.3)
    plan = sparta_b200.vbr_plan(v["rows"], 1000, 64, v["row_part"], v["nzcount"], v["jab"], 700, cta_pair=1,
                                acc_cols=256)
    segs, srows, chunks = plan["segs"], plan["srows"], plan["chunks"]
    assert np.all(segs["h_pad"] % 16 == 0) and np.all(segs["h"] <= segs["h_pad"]) and np.all(segs["h"] > 0)
    assert segs["h"].sum() == v["rows"]
    for sr in srows:
        s = segs[sr["seg_begin"]:sr["seg_begin"] + sr["seg_count"]]
        assert sr["n_cols"] <= 256 and sr["seg_count"] <= 32
        assert np.array_equal(s["tmem_col"], np.concatenate([[0], np.cumsum(s["h_pad"])[:-1]]))
        ch = chunks[sr["chunk_begin"]:sr["chunk_begin"] + sr["chunk_count"]]
        assert np.all(np.diff(ch["k0"]) > 0), "column blocks must be walked in ascending order"
        assert np.all(ch["mask"] != 0) and np.all(ch["mask"] < (1 << sr["seg_count"]))
    assert np.all(chunks["a_bytes"] % 2048 == 0) and np.all(chunks["ksteps"] >= 1) and np.all(chunks["ksteps"] <= 4)
    # every (super-row, column tile) exactly once
    items = plan["items"]
    assert len(items) == len(srows) * ((700 + 127) // 128)
    assert sorted(plan["cta_items"].tolist()) == list(range(len(items)))
    assert plan["stats"]["sched_imbalance"] < 1.5


def test_plan_shard_range(lib, oracle):
    rng = np.random.default_rng(8)
    v = random_vbr(rng, 12, 256, 32, [32] * 12, 0.5)
    n = 64
    Bm = rng.integers(-2, 3, size=(n, 256)).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n)
    cuts = sparta_b200.partition_block_rows(v["row_part"], v["nzcount"], 3)
    pieces = []
    for i in range(3):
        lo, hi = int(cuts[i]), int(cuts[i + 1])
        plan = sparta_b200.vbr_plan(v["rows"], 256, 32, v["row_part"], v["nzcount"], v["jab"], n,
                                    block_row_begin=lo, block_row_end=hi, cta_pair=1 + i % 2)
        rows_i = int(v["row_part"][hi] - v["row_part"][lo])
        # the shard's source offsets are relative to its first block
        mab_lo = int(sum(v["nzcount"][b] * 32 * 32 for b in range(lo)))
        pieces.append(sched_interp.run_plan(plan, v["mab"][mab_lo:], Bm, 256, n, rows_i))
    assert np.array_equal(np.concatenate(pieces, axis=1), Cref)


def test_empty_shard_is_empty(lib):
    """A balanced partition may hand a rank nothing (cuts[1] == 0 when the first block-row outweighs
    a whole share): an explicit [0, 0) range is an EMPTY shard, not the whole matrix; a zeroed
    options struct (no range given) still selects everything."""
    rng = np.random.default_rng(9)
    heights = [256, 8, 8, 8]
    v = random_vbr(rng, len(heights), 512, 32, heights, 1.0)
    cuts = sparta_b200.partition_block_rows(v["row_part"], v["nzcount"], 4)
    assert cuts[0] == 0 and cuts[-1] == len(heights)
    total_rows = 0
    for i in range(4):
        lo, hi = int(cuts[i]), int(cuts[i + 1])
        plan = sparta_b200.vbr_plan(v["rows"], 512, 32, v["row_part"], v["nzcount"], v["jab"], 64,
                                    block_row_begin=lo, block_row_end=hi)
        assert plan["stats"]["rows"] == int(v["row_part"][hi] - v["row_part"][lo])
        assert plan["stats"]["block_rows"] == hi - lo
        total_rows += plan["stats"]["rows"]
    assert total_rows == v["rows"]
    empty = sparta_b200.vbr_plan(v["rows"], 512, 32, v["row_part"], v["nzcount"], v["jab"], 64,
                                 block_row_begin=0, block_row_end=0)
    assert empty["stats"]["rows"] == 0 and empty["stats"]["nz_blocks"] == 0 and len(empty["items"]) == 0
    whole = sparta_b200.vbr_plan(v["rows"], 512, 32, v["row_part"], v["nzcount"], v["jab"], 64)
    assert whole["stats"]["rows"] == v["rows"]


def _longest_chain(chunks, srows, it):
    """MMAs the pass `it` issues into the busiest accumulator: a member's accumulator only takes
    the MMAs of the chunks it is present in (cut_passes, csrc/schedule.cpp)."""
    c0 = int(srows[it["srow"]]["chunk_begin"] + it["chunk_off"])
    cnt = int(it["count"]) & sparta_b200.lib.ITEM_COUNT_MASK
    ch = chunks[c0:c0 + cnt]
    per_member = [int(ch["ksteps"][(ch["mask"] >> m) & 1 == 1].sum()) for m in range(32)]
    return max(per_member) if cnt else 0


@pytest.mark.parametrize("pair", [1, 2], ids=["single", "pair"])
@pytest.mark.parametrize("precision,esize,max_chain", [("tf32", 4, 8), ("tf32", 4, 24), ("bf16", 2, 16), ("tf32", 4, 0)])
def test_bounded_accumulation_chains(oracle, lib, precision, esize, max_chain, pair):
    """max_chain cuts long super-rows into passes folded through the master accumulators; the
    interpreter checks pass order, the 256-column limit and that the product is unchanged."""
    rng = np.random.default_rng(21)
    heights = [64, 30, 64, 64, 7, 64, 64, 64, 16, 64]
    v = random_vbr(rng, len(heights), 2048, 64, heights, 0.6, values="int")
    n = 200
    Bm = rng.integers(-3, 4, size=(n, 2048)).astype(np.float32)
    plan = sparta_b200.vbr_plan(v["rows"], 2048, 64, v["row_part"], v["nzcount"], v["jab"], n,
                                precision=precision, cta_pair=pair, max_chain=max_chain)
    items, srows, chunks = plan["items"], plan["srows"], plan["chunks"]
    count = items["count"] & sparta_b200.lib.ITEM_COUNT_MASK
    limit = max_chain if max_chain else 256          # the tf32 default
    multi = (items["count"] & (sparta_b200.lib.ITEM_NOT_FIRST | sparta_b200.lib.ITEM_NOT_LAST)) != 0
    if max_chain:
        assert multi.any(), "the case is meant to need several passes"
        assert np.all(srows["n_cols"] <= 256)
    else:
        assert not multi.any()                       # 32 blocks x 8 MMAs = 256: fits the default
    for it in items:
        assert _longest_chain(chunks, srows, it) <= limit
    assert count.sum() == sum(int(sr["chunk_count"]) for sr in srows) * len(np.unique(items["j0"]))
    Cm = sched_interp.run_plan(plan, v["mab"], Bm, 2048, n, v["rows"], esize=esize)
    assert np.array_equal(Cm, oracle.vbr_multiply(v, Bm, n))


def test_unbounded_chain_keeps_wide_super_rows(lib):
    rng = np.random.default_rng(22)
    v = random_vbr(rng, 16, 4096, 64, [64] * 16, 0.9, values="int")
    a = sparta_b200.vbr_plan(v["rows"], 4096, 64, v["row_part"], v["nzcount"], v["jab"], 256, precision="bf16")
    b = sparta_b200.vbr_plan(v["rows"], 4096, 64, v["row_part"], v["nzcount"], v["jab"], 256, precision="tf32",
                             max_chain=-1)
    c = sparta_b200.vbr_plan(v["rows"], 4096, 64, v["row_part"], v["nzcount"], v["jab"], 256, precision="tf32")
    assert a["srows"]["n_cols"].max() == 512 and b["srows"]["n_cols"].max() == 512
    assert c["srows"]["n_cols"].max() == 256 and len(c["items"]) > len(c["srows"])


@pytest.mark.parametrize("pair", [1, 2], ids=["single", "pair"])
@pytest.mark.parametrize("precision,esize,max_chain", [("bf16", 2, 0), ("tf32", 4, 24)])
def test_split_pieces_cover_every_chunk_once(oracle, lib, precision, esize, max_chain, pair):
    """split_k = 2: the chunk lists are cut into one piece per worker; the interpreter checks that
    the pieces of every (super-row, tile) tile its chunk list, that only whole lists use plain
    stores, that every tile a piece adds into is zeroed first, and that the product is unchanged."""
    rng = np.random.default_rng(31)
    heights = [64, 64, 64, 30, 64, 64, 64, 64, 64, 17, 64, 64]
    v = random_vbr(rng, len(heights), 4096, 64, heights, 0.7, values="int")
    n = 300
    Bm = rng.integers(-3, 4, size=(n, 4096)).astype(np.float32)
    plan = sparta_b200.vbr_plan(v["rows"], 4096, 64, v["row_part"], v["nzcount"], v["jab"], n,
                                precision=precision, cta_pair=pair, max_chain=max_chain, split_k=2, num_ctas=20)
    st = plan["stats"]
    assert st["split_pieces"] > 0 and st["zero_tiles"] > 0
    assert st["sched_imbalance"] < 1.6   # forced on a tiny case: pieces have a minimum length
    atomic = (plan["items"]["count"] & sparta_b200.lib.ITEM_ATOMIC) != 0
    assert atomic.sum() == st["split_pieces"]
    if max_chain:
        for it in plan["items"]:
            assert _longest_chain(plan["chunks"], plan["srows"], it) <= max_chain
    Cm = sched_interp.run_plan(plan, v["mab"], Bm, 4096, n, v["rows"], esize=esize)
    assert np.array_equal(Cm, oracle.vbr_multiply(v, Bm, n))


@pytest.mark.parametrize("tiles,pair", [(2, 2), (4, 1)], ids=["2-tiles-pair", "4-tiles-single"])
def test_wide_items_split_plan(oracle, lib, tiles, pair):
    """Wide items (2 / 4 column tiles per item) under a forced split plan: a zero job covers the whole wide tile,
    every piece of a (super-row, wide tile) is drained once per tile, the product is unchanged; n leaves the last
    wide tile partly (and for 4 tiles mostly) beyond B."""
    rng = np.random.default_rng(33)
    heights = [64, 64, 30, 64, 17, 64]
    v = random_vbr(rng, len(heights), 4096, 64, heights, 0.5, values="int")
    n = 600
    Bm = rng.integers(-3, 4, size=(n, 4096)).astype(np.float32)
    plan = sparta_b200.vbr_plan(v["rows"], 4096, 64, v["row_part"], v["nzcount"], v["jab"], n, precision="bf16",
                                cta_pair=pair, split_k=2, num_ctas=24, wide_tiles=tiles)
    st = plan["stats"]
    assert st["wide_tiles"] == tiles and st["split_pieces"] > 0 and st["zero_tiles"] > 0
    assert int(plan["srows"]["n_cols"].max()) <= 512 // tiles
    Cm = sched_interp.run_plan(plan, v["mab"], Bm, 4096, n, v["rows"], esize=2)
    assert not np.isnan(Cm).any()
    assert np.array_equal(Cm, oracle.vbr_multiply(v, Bm, n))


def test_split_is_chosen_only_when_it_pays(lib):
    rng = np.random.default_rng(32)
    # few heavy super-rows for many workers: whole units cannot fill the grid
    v = random_vbr(rng, 24, 8192, 64, [64] * 24, 0.8, values="int")
    few = sparta_b200.vbr_plan(v["rows"], 8192, 64, v["row_part"], v["nzcount"], v["jab"], 2048, wide_tiles=1)
    never = sparta_b200.vbr_plan(v["rows"], 8192, 64, v["row_part"], v["nzcount"], v["jab"], 2048, split_k=1, wide_tiles=1)
    assert few["stats"]["split_pieces"] > 0 and never["stats"]["split_pieces"] == 0
    # whole units keep 24 of the 74 CTA pairs busy; the split plan fills the grid evenly (9 teams of
    # 8 column tiles and the 2 leftover pairs as a narrow tenth team)
    assert never["stats"]["grid"] == 48 and few["stats"]["grid"] == 148 and few["stats"]["team"] == 8
    assert few["stats"]["sched_imbalance"] < 1.25
    # (the library's own choice at n = 2048 is two column tiles per item: 6 super-rows x 4 wide tiles, split the same way)
    wide = sparta_b200.vbr_plan(v["rows"], 8192, 64, v["row_part"], v["nzcount"], v["jab"], 2048)
    assert wide["stats"]["wide_tiles"] == 2 and wide["stats"]["split_pieces"] > 0 and wide["stats"]["grid"] == 148
    # many units per worker: list scheduling is already balanced, nothing is split
    v = random_vbr(rng, 400, 1024, 64, [64] * 400, 0.3, values="int")
    many = sparta_b200.vbr_plan(v["rows"], 1024, 64, v["row_part"], v["nzcount"], v["jab"], 2048, num_ctas=16)
    assert many["stats"]["split_pieces"] == 0 and many["stats"]["zero_tiles"] == 0


BA_CASES = [
    # block_rows, cols, w, heights, density, m (rows of B)
    (6, 96, 16, [16] * 6, 0.5, 40),
    (5, 70, 16, [3, 17, 1, 64, 30], 0.6, 130),      # ragged heights become ragged k extents; cols % w != 0
    (4, 64, 3, [4, 3, 1, 1], 0.7, 2),
    (3, 300, 100, [64, 64, 20], 0.8, 16),
    (3, 64, 32, [200, 70, 9], 1.0, 8),              # tall block-rows: several K slabs per block
]


@pytest.mark.parametrize("pair", [1, 2], ids=["single", "pair"])
@pytest.mark.parametrize("precision,esize", [("bf16", 2), ("tf32", 4)])
@pytest.mark.parametrize("case", range(len(BA_CASES)))
def test_inverted_product_plan_matches_oracle(oracle, lib, case, precision, esize, pair):
    """C = B*A (-M 6): the schedule of the transposed operand, interpreted on the CPU, against the
    oracle's restatement of the arithmetic cublas_blockmat_multiplyBA intends and against numpy."""
    block_rows, cols, w, heights, density, m = BA_CASES[case]
    rng = np.random.default_rng(300 + case)
    v = random_vbr(rng, block_rows, cols, w, heights, density, values="int")
    Bt = rng.integers(-3, 4, size=(v["rows"], m)).astype(np.float32)      # row k = column k of B
    plan = sparta_b200.vbr_plan(v["rows"], cols, w, v["row_part"], v["nzcount"], v["jab"], m,
                                transposed=True, precision=precision, cta_pair=pair)
    st = plan["stats"]
    assert st["rows"] == cols and st["nz_blocks"] == v["jab"].size
    # the interpreter takes the dense operand as [n, K] with K = A's rows here
    Cm = sched_interp.run_plan(plan, v["mab"], np.ascontiguousarray(Bt.T), v["rows"], m, cols, esize=esize)
    Cref = oracle.vbr_multiply_BA(v, Bt, m)                                # [cols, m]
    assert not np.isnan(Cm).any()
    assert np.array_equal(Cm, Cref.T)
    from tests.util import vbr_to_dense
    assert np.array_equal(Cref, (Bt.T.astype(np.float64) @ vbr_to_dense(v)).T.astype(np.float32))


def test_modelled_partition_is_contiguous_and_no_worse_than_area(lib):
    """sparta_partition_block_rows_modelled: cuts are monotone, cover every block-row, and the
    slowest shard's modelled time is not above the area partition's."""
    rng = np.random.default_rng(71)
    heights = [64] * 120
    # dense head, sparse tail: equal areas are not equal times
    dens = np.concatenate([np.full(30, 0.9), np.full(90, 0.08)])
    rows = 0
    row_part, nzcount, jab = [0], [], []
    for b, h in enumerate(heights):
        cols = np.flatnonzero(rng.random(128) < dens[b])
        nzcount.append(len(cols)); jab.extend(cols.tolist()); rows += h; row_part.append(rows)
    row_part, nzcount, jab = np.array(row_part), np.array(nzcount), np.array(jab)
    n = 1024
    for parts in (2, 4):
        area = sparta_b200.partition_block_rows(row_part, nzcount, parts)
        model = sparta_b200.partition_block_rows_modelled(rows, 128 * 64, 64, row_part, nzcount, jab, n, parts)
        assert model[0] == 0 and model[-1] == len(nzcount) and np.all(np.diff(model) >= 0)

        def worst(cuts):
            t = []
            for i in range(parts):
                p = sparta_b200.vbr_plan(rows, 128 * 64, 64, row_part, nzcount, jab, n,
                                         block_row_begin=int(cuts[i]), block_row_end=int(cuts[i + 1]))
                t.append(p["stats"]["sched_max_cycles"])
            return max(t)
        assert worst(model) <= worst(area) * 1.001


@pytest.mark.parametrize("pair", [1, 2], ids=["single", "pair"])
@pytest.mark.parametrize("precision,esize", [("bf16", 2), ("tf32", 4)])
def test_fused_short_block_rows(oracle, lib, precision, esize, pair):
    """Runs of consecutive short block-rows share one 16-row segment with the union of their
    column-block lists (fuse_short_block_rows): same product, same nztot / block count as the
    original VBR, fewer segments and no more image bytes than the unfused plan."""
    rng = np.random.default_rng(81)
    heights = [1, 1, 2, 1, 5, 1, 1, 1, 1, 3, 64, 1, 1, 1, 9, 8, 1, 1, 40, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2]
    v = random_vbr(rng, len(heights), 640, 32, heights, 0.35, values="int")
    n = 70
    Bm = rng.integers(-3, 4, size=(n, 640)).astype(np.float32)
    fused = sparta_b200.vbr_plan(v["rows"], 640, 32, v["row_part"], v["nzcount"], v["jab"], n,
                                 precision=precision, cta_pair=pair)
    plain = sparta_b200.vbr_plan(v["rows"], 640, 32, v["row_part"], v["nzcount"], v["jab"], n,
                                 precision=precision, cta_pair=pair, fuse_rows=1)
    for p in (fused, plain):
        assert p["stats"]["nztot"] == v["mab"].size and p["stats"]["nz_blocks"] == v["jab"].size
        assert p["segs"]["h"].sum() == v["rows"]
    assert len(fused["segs"]) < len(plain["segs"]) == len(heights)        # every block-row here fits one segment
    assert fused["stats"]["a_packed_bytes"] <= plain["stats"]["a_packed_bytes"]
    assert fused["stats"]["chunks"] < plain["stats"]["chunks"]
    assert np.any(fused["jobs"]["r_base"] > 0)
    Cref = oracle.vbr_multiply(v, Bm, n)
    for p in (fused, plain):
        Cm = sched_interp.run_plan(p, v["mab"], Bm, 640, n, v["rows"], esize=esize)
        assert not np.isnan(Cm).any() and np.array_equal(Cm, Cref)


@pytest.mark.parametrize("n,num_ctas,expect", [(1024, 12, (4, 2)), (700, 16, (3, 1)), (1024, 20, (4, 2))])
def test_leftover_workers_form_a_narrow_team(oracle, lib, n, num_ctas, expect):
    """Whole-unit plans: the CTA pairs a multiple of the team width leaves over work as one
    narrower team whose members take a group's column tiles in turns.  Every (super-row, tile) is
    still covered exactly once (checked by the interpreter) and the narrow team gets less work."""
    rng = np.random.default_rng(95)
    v = random_vbr(rng, 48, 1024, 64, [64] * 48, 0.5, values="int")
    Bm = rng.integers(-3, 4, size=(n, 1024)).astype(np.float32)
    plan = sparta_b200.vbr_plan(v["rows"], 1024, 64, v["row_part"], v["nzcount"], v["jab"], n,
                                num_ctas=num_ctas, split_k=1)
    team, spare = expect
    workers = num_ctas // 2
    assert plan["stats"]["team"] == team and plan["stats"]["grid"] == 2 * (workers // team * team + spare)
    items_per_worker = np.diff(plan["cta_ptr"])
    assert items_per_worker[-1] > 0                                 # the narrow team has work ...
    cnt = plan["items"]["count"] & sparta_b200.lib.ITEM_COUNT_MASK
    load = [int(cnt[plan["cta_items"][plan["cta_ptr"][w]:plan["cta_ptr"][w + 1]]].sum()) for w in range(len(items_per_worker))]
    assert max(load) <= 1.35 * np.mean(load)                         # ... and is not the straggler
    Cm = sched_interp.run_plan(plan, v["mab"], Bm, 1024, n, v["rows"])
    assert np.array_equal(Cm, oracle.vbr_multiply(v, Bm, n))


def test_measured_partition_moves_cuts_towards_slow_shards(lib):
    """sparta_partition_block_rows_measured: equal measured / modelled ratios reproduce the modelled
    cuts; a shard that ran slower than modelled gives block-rows away."""
    rng = np.random.default_rng(72)
    row_part = np.arange(0, 64 * 201, 64)
    nzcount, jab = [], []
    for b in range(200):
        cols = np.flatnonzero(rng.random(256) < 0.3)
        nzcount.append(len(cols)); jab.extend(cols.tolist())
    nzcount, jab = np.array(nzcount), np.array(jab)
    rows, n = int(row_part[-1]), 1024
    base = sparta_b200.partition_block_rows_modelled(rows, 256 * 64, 64, row_part, nzcount, jab, n, 4)
    model = []
    for i in range(4):
        p = sparta_b200.vbr_plan(rows, 256 * 64, 64, row_part, nzcount, jab, n, block_row_begin=int(base[i]),
                                 block_row_end=int(base[i + 1]))
        model.append(p["stats"]["sched_max_cycles"])
    same = sparta_b200.partition_block_rows_measured(rows, 256 * 64, 64, row_part, nzcount, jab, n, 4, base,
                                                     [m * 1e-6 for m in model], model)
    assert np.array_equal(same, base)
    slow_first = sparta_b200.partition_block_rows_measured(rows, 256 * 64, 64, row_part, nzcount, jab, n, 4, base,
                                                           [model[0] * 1.3e-6] + [m * 1e-6 for m in model[1:]], model)
    assert slow_first[1] < base[1] and slow_first[0] == 0 and slow_first[-1] == 200
