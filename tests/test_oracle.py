"""The oracle (oracle/sparta_oracle.cpp) pinned against the reference: golden vectors captured
from the compiled reference, and -- where oracle/_ref exists -- the unmodified reference itself
on randomised inputs.  CPU only."""
import json
import os

import numpy as np
import pytest

from sparta_b200 import synth
from tests.util import random_vbr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
with open(os.path.join(GOLDEN, "reference_vectors.json")) as f:
    VECTORS = json.load(f)

ARRAY_KEYS = ["csr_rowptr", "csr_colind", "csr_val", "grouping", "row_part", "nzcount", "jab", "mab"]
SCALAR_KEYS = ["csr_rows", "csr_cols", "csr_nnz", "comparison_counter", "merge_counter",
               "VBR_nzcount", "VBR_nzblocks_count", "VBR_longest_row", "rows", "cols", "block_rows",
               "block_cols", "block_col_size", "nztot"]
FLOAT_KEYS = ["average_merge_tau", "average_row_distance", "VBR_average_height"]


def _same(res, rec):
    for k in ARRAY_KEYS:
        assert np.array_equal(res[k], np.asarray(rec[k], dtype=res[k].dtype)), k
    for k in SCALAR_KEYS:
        assert res[k] == rec[k], k
    for k in FLOAT_KEYS:
        if rec[k] is None:
            assert res[k] != res[k], k
        else:
            assert np.float32(res[k]) == np.float32(rec[k]), k


@pytest.mark.parametrize("idx", range(len(VECTORS)))
def test_oracle_matches_golden_structure_and_product(oracle, idx):
    rec = VECTORS[idx]
    res = oracle.run(os.path.join(GOLDEN, rec["file"]), **rec["flags"])
    _same(res, rec)
    n = rec["n"]
    Cm = oracle.vbr_multiply(res, np.asarray(rec["B"], dtype=np.float32), n)
    assert np.array_equal(Cm.reshape(-1), np.asarray(rec["C"], dtype=np.float32))


def test_survey_vectors_test_matrix(oracle):
    """The numbers printed in SURVEY.md 8(c) for data/TEST_matrix_weighted.el -b 3 -t 0.6."""
    res = oracle.run(os.path.join(GOLDEN, "TEST_matrix_weighted.el"), b=3, t=0.6)
    assert res["grouping"].tolist() == [0, 1, 1, 1, 0, 5, 0, 0, 8]
    assert res["row_part"].tolist() == [0, 4, 7, 8, 9]
    assert res["nzcount"].tolist() == [0, 3, 1, 1]
    assert res["jab"].tolist() == [0, 1, 2, 2, 0]
    assert res["mab"].astype(int).tolist() == [0, 0, 0, 0, 0, 1, 5, 0, 0, 0, 0, 1, 0, 0, 0, 8, 1, 0, 0,
                                               1, 0, 0, 0, 3, 7, 1, 8, 2, 0, 0, 0, 5, 0]
    assert (res["VBR_nzcount"], res["VBR_nzblocks_count"], res["VBR_longest_row"]) == (33, 5, 3)
    assert (res["merge_counter"], res["comparison_counter"]) == (5, 13)
    B = (np.arange(18) + 1).astype(np.float32)
    Cm = oracle.vbr_multiply(res, B, 2)
    assert Cm.tolist() == [[0, 0, 0, 0, 126, 22, 102, 14, 10], [0, 0, 0, 0, 306, 49, 219, 32, 55]]


def test_similarities_vectors(oracle):
    """test/general/TEST_similarities.cpp:14-36 values probed in SURVEY 8(c)."""
    a, b = [1, 2, 5, 10, 12, 20], [0, 2, 4, 10, 16]
    assert oracle.distance(0, a, 1, b, 1, 3) == 3.0
    assert oracle.distance(1, a, 1, b, 1, 3) == 0.5
    assert oracle.distance(0, a, 1, b, 1, 1) == 7.0
    assert abs(oracle.distance(1, a, 1, b, 1, 1) - 0.777778) < 1e-6
    assert oracle.distance(2, a, 1, b, 1, 3) == 3.0 and oracle.distance(3, a, 1, b, 1, 3) == 0.5


def test_csr_vs_vbr_multiply_equal(oracle):
    """test/general/TEST_matrices.cpp:9-57: fixed blocking, B = ones with 5 columns; the CSR and
    VBR products must be identical (un-permuted rows because the blocking is fixed-size)."""
    res = oracle.run(os.path.join(GOLDEN, "TEST_matrix_weighted.el"), a=2, b=3, B=3)
    B = np.ones(5 * 9, dtype=np.float32)
    Cv = oracle.vbr_multiply(res, B, 5)
    Cc = oracle.csr_multiply(res["csr_rows"], res["csr_rowptr"], res["csr_colind"], res["csr_val"], 0, B, 5)
    assert np.array_equal(Cv, Cc)
    assert Cv[0].tolist() == [0, 20, 3, 13, 0, 2, 0, 0, 5]


def test_merge_rows_is_not_a_union(oracle):
    """utilities.cpp:145-173 drops the tail of A above the last B entry inside A's range."""
    assert oracle.merge_rows([1, 5, 9], [3]).tolist() == [1, 3]
    assert oracle.merge_rows([1, 5], [7]).tolist() == [7]
    assert oracle.merge_rows([1, 5, 9], []).tolist() == []
    assert oracle.merge_rows([1, 5, 9], [5, 20]).tolist() == [1, 5, 20]


def test_bellpack_repack(oracle):
    res = oracle.run(os.path.join(GOLDEN, "rmat8.el"), P=1, a=2, b=16, B=16, F=1)
    bs, ind, vals = oracle.bellpack_from_vbr(res)
    assert bs == 16 and ind.shape[0] == res["rows"] // 16 and ind.shape[1] == res["nzcount"].max()
    from sparta_b200.api import VBR, bellpack_from_vbr
    v = VBR(res["rows"], res["cols"], 16, res["row_part"], res["nzcount"], res["jab"], res["mab"])
    bs2, ind2, vals2 = bellpack_from_vbr(v)
    assert bs2 == bs and np.array_equal(ind, ind2) and np.array_equal(vals, vals2)


FLAG_SETS = [
    dict(a=3, b=8, t=0.6), dict(a=4, b=8, t=0.6), dict(a=5, b=8, B=8, t=0.6),
    dict(a=5, b=16, B=32, t=0.3), dict(a=5, b=8, B=7, t=0.9), dict(a=2, b=8, B=8, F=1),
    dict(a=3, b=8, B=8, t=0.5, F=1), dict(a=3, b=4, t=0.4, m=0), dict(a=5, b=8, B=8, t=0.6, r=2, s=7),
    dict(a=3, b=8, t=0.6, r=-1), dict(a=4, b=8, t=1.0), dict(a=4, b=8, t=0.0), dict(a=6, b=8),
    dict(a=5, b=8, B=8, t=0.6, g=1), dict(a=3, b=8, t=0.6, g=1, p=0), dict(a=5, b=64, B=64, t=0.6),
]


@pytest.mark.parametrize("flags", FLAG_SETS, ids=lambda f: "-".join(f"{k}{v}" for k, v in f.items()))
@pytest.mark.parametrize("kind", ["rmat", "er_weighted"])
def test_oracle_equals_reference_build(oracle, reference, tmp_path, kind, flags):
    """Bit-for-bit against the unmodified reference sources (skipped where oracle/_ref is absent)."""
    if kind == "rmat":
        r, c = synth.rmat_edges(9, 6000, seed=1)
        r, c = synth.pin_shape(r, c, 512, 512)
        path = str(tmp_path / "m.el")
        synth.write_el(path, r, c)
        flags = dict(flags, P=1)
    else:
        rng = np.random.default_rng(5)
        r, c = synth.er_edges(300, 280, 0.02, seed=2)
        path = str(tmp_path / "m.el")
        synth.write_el(path, r, c, vals=rng.uniform(-1, 1, len(r)))
    ro = oracle.run(path, **flags)
    rr = reference.run(path, **flags)
    for k, a in ro.items():
        b = rr[k]
        if isinstance(a, np.ndarray):
            assert np.array_equal(a, b), k
        else:
            assert a == b or (a != a and b != b), k
    rng = np.random.default_rng(0)
    # VBR::multiply reads B[w*jb + k + j*cols] for k < w even in the last, ragged column block
    # (vbr.cpp:362), i.e. up to w-1 floats past the end of the last column of B; the matching A
    # entries are zero.  Give the buffer that much finite slack so 0 * garbage cannot make NaNs.
    cols, w = int(rr["cols"]), int(rr["block_col_size"])
    B = np.zeros(3 * cols + w, dtype=np.float32)
    B[:3 * cols] = rng.uniform(0, 1, size=3 * cols)
    assert np.array_equal(oracle.vbr_multiply(rr, B, 3), reference.vbr_multiply(rr, B, 3))


def test_BA_reference_loop_literal_vs_intended(oracle):
    """Pins the inverted product C = B*A (-M 6).  The reference has no CPU routine for it; its GPU
    loop (cublas_blockmat_multiplyBA, src/cuda/cuda_utilities.cpp:640-690) is restated index for
    index in oracle_ref_multiplyBA_literal (the GEMMs done in fp32 on the CPU).  Where that loop's
    B offset `d_B + block_col_size*ib` (:645) addresses the block of B it is meant to multiply --
    B_rows == 1 with constant heights equal to the column-block width (-b == -B), or a single
    block-row -- the literal loop and the product the library computes (oracle_vbr_multiply_BA) are
    the SAME numbers; elsewhere the literal loop is the intended product of a row-shifted B, which
    is why the library implements the intended arithmetic (include/sparta_b200.h)."""
    rng = np.random.default_rng(33)
    # (1) B_rows = 1, h == w, several block-rows, every block-row non-empty
    h = w = 8
    v = random_vbr(rng, 6, 48, w, [h] * 6, 0.6, values="int", empty_rows=False)
    Bt = rng.integers(-3, 4, size=(v["rows"], 1)).astype(np.float32)
    lit, oor = oracle.ref_multiplyBA_literal(v, Bt, 1)
    assert not oor
    assert np.array_equal(lit[:v["cols"]], oracle.vbr_multiply_BA(v, Bt, 1))
    # (2) one block-row, any B_rows and any height
    v1 = random_vbr(rng, 1, 40, 8, [5], 0.8, values="int", empty_rows=False)
    Bt1 = rng.integers(-3, 4, size=(5, 7)).astype(np.float32)
    lit1, oor1 = oracle.ref_multiplyBA_literal(v1, Bt1, 7)
    assert not oor1 and np.array_equal(lit1[:40], oracle.vbr_multiply_BA(v1, Bt1, 7))
    # (3) B_rows = 3: the literal loop reads element w*ib + i + k*B_rows of B, i.e. it multiplies
    # block-row ib by rows shifted w*ib down inside the first h columns.  It equals the intended
    # product of the matrix B' whose column row_part[ib] + k holds exactly those elements.
    m = 3
    Bt3 = rng.integers(-3, 4, size=(v["rows"], m)).astype(np.float32)
    lit3, oor3 = oracle.ref_multiplyBA_literal(v, Bt3, m)
    assert not oor3
    intended = oracle.vbr_multiply_BA(v, Bt3, m)
    assert not np.array_equal(lit3[:v["cols"]], intended)
    flat = Bt3.reshape(-1)     # column-major m x rows: element (i, k) at i + k*m
    shifted = np.zeros_like(Bt3)
    for ib in range(6):
        for k in range(h):
            for i in range(m):
                shifted[h * ib + k, i] = flat[w * ib + i + k * m]
    assert np.array_equal(lit3[:v["cols"]], oracle.vbr_multiply_BA(v, shifted, m))
    # (4) variable heights: the loop keeps using the first block-row's height (:637) and strides mab by
    # it (:685), so for heights that differ it walks mab out of step -- different from the intended product
    vv = random_vbr(rng, 3, 24, 8, [8, 4, 8], 1.0, values="int", empty_rows=False)
    Btv = rng.integers(-3, 4, size=(vv["rows"], 1)).astype(np.float32)
    litv, _ = oracle.ref_multiplyBA_literal(vv, np.concatenate([Btv, np.zeros((16, 1), np.float32)]), 1)
    assert not np.array_equal(litv[:24], oracle.vbr_multiply_BA(vv, Btv, 1))
