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
        c0 = srows[it["srow"]]["chunk_begin"] + it["chunk_off"]
        n_mma = int(chunks["ksteps"][c0:c0 + (int(it["count"]) & sparta_b200.lib.ITEM_COUNT_MASK)].sum())
        assert n_mma <= limit
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
