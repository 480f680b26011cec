"""Parity tests proper: the sm_100a kernel, called through the C ABI, against the oracle
(oracle/sparta_oracle.cpp restating VBR::multiply, src/general/vbr.cpp:323-372).

Tolerances (norm = max|C - Cref| / max|Cref|):
  * integer-valued operands (exact in bf16/fp16/tf32, sums exact in fp32): bit-exact;
  * general operands vs the oracle fed THE SAME fp32 inputs: <= 2e-2 for bf16/fp16;
  * general operands vs the oracle fed the operands rounded to the kernel's input precision
    (only the fp32 accumulation order differs): <= 1e-5 for tf32, bf16 and fp16.
"""
import json
import os

import numpy as np
import pytest

import sparta_b200
from sparta_b200 import synth
from sparta_b200.api import VBR, bellpack_from_vbr, bellpack_spmm, vbr_spmm
from tests.util import random_vbr, rel_err, round_to, vbr_to_dense

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL_ROUNDED = 1e-5
TOL_UNROUNDED = {"bf16": 2e-2, "fp16": 2e-2, "tf32": 2e-3}


def gpu_multiply(v, Bm, n, precision="bf16", **opts):
    """C as [n, rows] through the handle API."""
    h = sparta_b200.Handle.from_vbr(v["rows"], v["cols"], v["block_col_size"], v["row_part"],
                                    v["nzcount"], v["jab"], v["mab"], precision=precision, **opts)
    try:
        h.set_B(Bm, Bm.shape[1], n)
        h.run()
        out = np.zeros((n, v["rows"]), dtype=np.float32)
        h.get_C(out, v["rows"])
        st = h.stats()   # one launch of the tensor-core kernel if it has work items, one of the gather kernel if it has rows
        assert st["kernel_launches"] == (1 if st["items"] else 0) + (1 if st["gather_nnz"] else 0)
        return out
    finally:
        h.close()


def rounded(v, precision):
    w = dict(v)
    w["mab"] = round_to(v["mab"], precision)
    return w


CASES = [
    (6, 96, 16, [16] * 6, 0.5, 40, {}),
    (5, 70, 16, [3, 17, 1, 64, 30], 0.6, 130, {}),
    (4, 64, 3, [4, 3, 1, 1], 0.7, 2, {}),
    (3, 300, 100, [64, 64, 20], 0.8, 16, {}),
    (7, 128, 64, [64] * 7, 0.4, 256, {"acc_cols": 256}),
    (2, 64, 32, [200, 70], 1.0, 8, {}),
    (40, 64, 8, [1] * 40, 0.3, 8, {"acc_cols": 512}),
    (3, 64, 16, [5, 7, 9], 0.0, 8, {}),                    # no nonzero block at all -> C = 0
    (24, 1024, 64, [64] * 24, 0.5, 384, {"acc_cols": 256}),  # several items per CTA, pipeline wraps
    (40, 2048, 64, [64] * 37 + [63, 30, 7], 0.5, 700, {"num_ctas": 8}),   # many items per worker, 512-col super-rows
    (9, 640, 64, [64, 63, 30, 64, 1, 128, 7, 64, 33], 0.7, 200, {"acc_cols": 512, "panel_stages": 3}),
    (16, 512, 64, [64] * 16, 1.0, 128, {"panel_stages": 2}),
    (3, 2048, 128, [256, 100, 16], 0.9, 64, {"seg_rows": 256, "acc_cols": 512}),
]


MODES = {"single": dict(cta_pair=1), "pair": dict(cta_pair=2), "pair_input_order": dict(cta_pair=2, row_order=1)}


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_integer_operands_bit_exact(oracle, lib, case, precision, mode):
    block_rows, cols, w, heights, density, n, opts = CASES[case]
    opts = dict(opts, **MODES[mode])
    rng = np.random.default_rng(100 + case)
    v = random_vbr(rng, block_rows, cols, w, heights, density, values="int")
    Bm = rng.integers(-3, 4, size=(n, cols)).astype(np.float32)
    Cg = gpu_multiply(v, Bm, n, precision, **opts)
    Cref = oracle.vbr_multiply(v, Bm, n)
    assert np.array_equal(Cg, Cref)


@pytest.mark.parametrize("mode", ["single", "pair"])
@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("case", [1, 3, 8, 9, 11])
def test_real_operands_within_tolerance(oracle, lib, case, precision, mode):
    block_rows, cols, w, heights, density, n, opts = CASES[case]
    opts = dict(opts, **MODES[mode])
    rng = np.random.default_rng(200 + case)
    v = random_vbr(rng, block_rows, cols, w, heights, density, values="uniform")
    Bm = rng.random((n, cols), dtype=np.float32)  # uniform(0,1) like cuda_multiply.cpp:36-44
    Cg = gpu_multiply(v, Bm, n, precision, **opts)
    C_same_inputs = oracle.vbr_multiply(v, Bm, n)
    C_rounded = oracle.vbr_multiply(rounded(v, precision), round_to(Bm, precision), n)
    assert rel_err(Cg, C_rounded) <= TOL_ROUNDED
    assert rel_err(Cg, C_same_inputs) <= TOL_UNROUNDED[precision]


with open(os.path.join(GOLDEN, "reference_vectors.json")) as f:
    VECTORS = json.load(f)


@pytest.mark.parametrize("idx", range(len(VECTORS)))
def test_golden_reference_products(lib, idx):
    """C of the compiled reference's own VBR::multiply on its own VBR arrays (tests/golden)."""
    rec = VECTORS[idx]
    v = {k: (np.asarray(rec[k]) if isinstance(rec[k], list) else rec[k])
         for k in ("rows", "cols", "block_col_size", "row_part", "nzcount", "jab", "mab")}
    n = rec["n"]
    Bm = np.asarray(rec["B"], dtype=np.float32).reshape(n, rec["cols"])
    Cref = np.asarray(rec["C"], dtype=np.float32).reshape(n, rec["rows"])
    Cg = gpu_multiply(v, Bm, n, "tf32")
    if rec["file"] == "er_weighted.el":   # weights have 3 decimals: not exact in tf32
        assert rel_err(Cg, Cref) <= TOL_UNROUNDED["tf32"]
    else:
        assert np.array_equal(Cg, Cref)
    Cg = gpu_multiply(v, Bm, n, "bf16")
    if rec["file"] == "er_weighted.el":
        assert rel_err(Cg, Cref) <= TOL_UNROUNDED["bf16"]
    else:
        assert np.array_equal(Cg, Cref)


def test_one_shot_reference_signature(oracle, lib):
    """sparta_vbr_spmm = upload + multiply + download, the data flow of
    cublas_fixed_blocks_multiply (cuda_utilities.cpp:39-209)."""
    rng = np.random.default_rng(5)
    v = random_vbr(rng, 10, 320, 32, [32] * 10, 0.5, values="int")
    Bm = rng.integers(-3, 4, size=(96, 320)).astype(np.float32)
    A = VBR(v["rows"], v["cols"], 32, v["row_part"], v["nzcount"], v["jab"], v["mab"])
    Cg, dt = vbr_spmm(A, Bm, 96, "bf16")
    assert dt > 0
    assert np.array_equal(Cg, oracle.vbr_multiply(v, Bm, 96))


def test_accumulate_beta_one_and_idempotence(oracle, lib):
    rng = np.random.default_rng(6)
    v = random_vbr(rng, 6, 256, 64, [64, 20, 64, 1, 64, 64], 0.6, values="int")
    n = 72
    Bm = rng.integers(-3, 4, size=(n, 256)).astype(np.float32)
    C0 = rng.integers(-5, 6, size=(n, v["rows"])).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n, C_init=C0)   # reference semantics: C += A*B
    h = sparta_b200.Handle.from_vbr(v["rows"], 256, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                    accumulate=1)
    h.set_B(Bm, 256, n)
    h.set_C(C0, v["rows"])
    h.run()
    out = np.zeros_like(C0)
    h.get_C(out, v["rows"])
    assert np.array_equal(out, Cref)
    h.close()
    # accumulate = 0: running twice gives the same C (no hidden state between launches)
    h = sparta_b200.Handle.from_vbr(v["rows"], 256, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"])
    h.set_B(Bm, 256, n)
    h.run()
    a = h.get_C(np.zeros_like(C0), v["rows"]).copy()
    h.run()
    h.run()
    b = h.get_C(np.zeros_like(C0), v["rows"])
    assert np.array_equal(a, b) and np.array_equal(a, oracle.vbr_multiply(v, Bm, n))
    # new B with a different n on the same handle
    B2 = rng.integers(-3, 4, size=(200, 256)).astype(np.float32)
    h.set_B(B2, 256, 200)
    h.run()
    c = h.get_C(np.zeros((200, v["rows"]), np.float32), v["rows"])
    assert np.array_equal(c, oracle.vbr_multiply(v, B2, 200))
    h.close()


@pytest.mark.parametrize("mode", ["single", "pair"])
def test_back_to_back_launches_overlap_safely(oracle, lib, mode):
    """Launches of one handle enqueued back to back are chained with programmatic dependent launch: the next
    grid's copies and MMAs may start while the previous grid's slowest CTAs finish, its epilogue waits for
    that grid (griddepcontrol.wait) before touching C.  Checked where an early write would show: split
    pieces (tiles zeroed in-kernel, then red.add), beta = 1 (C is read back by the next launch), and a set_B
    between two launches (B changes: the chain must be broken).  Integer operands: bit-exact."""
    rng = np.random.default_rng(47)
    heights = [64, 64, 64, 30, 64, 64, 64, 64, 64, 17, 64, 64, 1, 64]
    v = random_vbr(rng, len(heights), 4096, 64, heights, 0.7, values="int")
    n = 600
    Bm = rng.integers(-3, 4, size=(n, 4096)).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n)
    h = sparta_b200.Handle.from_vbr(v["rows"], 4096, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                    split_k=2, gather_max_height=-1, n_hint=n, **MODES[mode])
    try:
        h.set_B(Bm, 4096, n)
        assert h.stats()["split_pieces"] > 0
        for _ in range(8):
            h.run_async()
        h.synchronize()
        assert np.array_equal(h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"]), Cref)
        # a new B between two launches, nothing synchronised in between
        B2 = rng.integers(-3, 4, size=(n, 4096)).astype(np.float32)
        h.run_async()
        h.set_B(B2, 4096, n)
        h.run_async()
        h.run_async()
        h.synchronize()
        assert np.array_equal(h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"]), oracle.vbr_multiply(v, B2, n))
    finally:
        h.close()
    # beta = 1: five chained launches add the product five times
    C0 = rng.integers(-5, 6, size=(n, v["rows"])).astype(np.float32)
    h = sparta_b200.Handle.from_vbr(v["rows"], 4096, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                    accumulate=1, gather_max_height=-1, n_hint=n, **MODES[mode])
    try:
        h.set_B(Bm, 4096, n)
        h.set_C(C0, v["rows"])
        for _ in range(5):
            h.run_async()
        h.synchronize()
        out = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"])
    finally:
        h.close()
    assert np.array_equal(out, C0 + 5 * Cref)


def test_padded_leading_dimensions(oracle, lib):
    rng = np.random.default_rng(7)
    v = random_vbr(rng, 5, 100, 16, [16, 5, 16, 16, 11], 0.7, values="int")
    n, ldb, ldc = 33, 120, 80
    Bfull = rng.integers(-3, 4, size=(n, ldb)).astype(np.float32)
    h = sparta_b200.Handle.from_vbr(v["rows"], 100, 16, v["row_part"], v["nzcount"], v["jab"], v["mab"])
    h.set_B(Bfull, ldb, n)
    h.run()
    out = np.full((n, ldc), -7.0, dtype=np.float32)
    h.get_C(out, ldc)
    h.close()
    Cref = oracle.vbr_multiply(v, Bfull, n, ldb=ldb)
    assert np.array_equal(out[:, :v["rows"]], Cref)
    assert np.all(out[:, v["rows"]:] == -7.0)


def test_row_block_shards_reassemble(oracle, lib):
    """Multi-GPU partition on one device: every shard computes its own C slab."""
    rng = np.random.default_rng(8)
    heights = rng.integers(1, 100, size=30)
    v = random_vbr(rng, 30, 512, 64, heights, 0.4, values="int")
    n = 130
    Bm = rng.integers(-3, 4, size=(n, 512)).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n)
    cuts = sparta_b200.partition_block_rows(v["row_part"], v["nzcount"], 4)
    slabs = []
    for i in range(4):
        lo, hi = int(cuts[i]), int(cuts[i + 1])
        rows_i = int(v["row_part"][hi] - v["row_part"][lo])
        h = sparta_b200.Handle.from_vbr(v["rows"], 512, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                        block_row_begin=lo, block_row_end=hi)
        h.set_B(Bm, 512, n)
        h.run()
        slabs.append(h.get_C(np.zeros((n, rows_i), np.float32), rows_i) if rows_i else np.zeros((n, 0), np.float32))
        assert h.stats()["rows"] == rows_i
        h.close()
    assert np.array_equal(np.concatenate(slabs, axis=1), Cref)


def test_bellpack_path_row_major(oracle, lib):
    """-M 3 / -M 8 replacement: Blocked-ELL A, row-major B and C (cuda_utilities.cpp:1581-1591)."""
    res = oracle.run(os.path.join(GOLDEN, "rmat8.el"), P=1, a=2, b=16, B=16, F=1)
    A = VBR(res["rows"], res["cols"], 16, res["row_part"], res["nzcount"], res["jab"], res["mab"])
    bs, ind, vals = bellpack_from_vbr(A)
    bs_o, ind_o, vals_o = oracle.bellpack_from_vbr(res)
    assert bs == bs_o and np.array_equal(ind, ind_o) and np.array_equal(vals, vals_o)
    rng = np.random.default_rng(9)
    n = 50
    B_rm = rng.integers(-3, 4, size=(res["cols"], n)).astype(np.float32)   # row-major cols x n
    Cg, dt = bellpack_spmm(res["rows"], res["cols"], bs, ind, vals, B_rm, n, "bf16")
    Cref = oracle.vbr_multiply(res, np.ascontiguousarray(B_rm.T), n)       # [n, rows]
    assert np.array_equal(Cg, Cref.T)


def test_device_resident_B_and_C(oracle, lib):
    """B arrives as a device pointer (the NCCL-broadcast path) and C is read back on the device."""
    import torch
    rng = np.random.default_rng(10)
    v = random_vbr(rng, 8, 256, 64, [64] * 8, 0.5, values="int")
    n = 256
    Bm = rng.integers(-3, 4, size=(n, 256)).astype(np.float32)
    Bd = torch.from_numpy(Bm).cuda()
    h = sparta_b200.Handle.from_vbr(v["rows"], 256, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"])
    h.set_B_device(Bd.data_ptr(), 256, n)
    h.run()
    Cd = torch.zeros((n, v["rows"]), dtype=torch.float32, device="cuda")
    h.get_C_device(Cd.data_ptr(), v["rows"])
    torch.cuda.synchronize()
    assert np.array_equal(Cd.cpu().numpy(), oracle.vbr_multiply(v, Bm, n))
    h.close()


def test_config2_shape_er_bellpack_blocking(lib):
    """BASELINE config #2 at full size: ER 16384^2, fixed 64x64 blocks, ~10% block density,
    n = 1024, bf16.  Too big for the serial oracle, so: (i) sampled block-rows recomputed in
    fp64 from the bf16-rounded operands, (ii) linearity C(B1 + B2) = C(B1) + C(B2) on
    integer operands (exact)."""
    N, w, n = 16384, 64, 1024
    r, c = synth.er_edges(N, N, 2.57e-5, seed=1)
    r, c = synth.pin_shape(r, c, N, N)
    rowptr, colind, val = synth.csr_from_edges(r, c, N)
    from sparta_b200.lib import host_vbr_fill
    v = host_vbr_fill(N, N, rowptr, colind, None, np.arange(N) // w, w, w, True)
    assert v["block_rows"] == 256 and 5500 < len(v["jab"]) < 7500
    rng = np.random.default_rng(2)
    B1 = rng.integers(-3, 4, size=(n, N)).astype(np.float32)
    B2 = rng.integers(-3, 4, size=(n, N)).astype(np.float32)
    h = sparta_b200.Handle.from_vbr(N, N, w, v["row_part"], v["nzcount"], v["jab"], v["mab"])
    outs = []
    for Bm in (B1, B2, B1 + B2):
        h.set_B(Bm, N, n)
        h.run()
        outs.append(h.get_C(np.zeros((n, N), np.float32), N).copy())
    h.close()
    assert np.array_equal(outs[0] + outs[1], outs[2])
    A = vbr_to_dense({**v, "row_part": v["row_part"][:3], "nzcount": v["nzcount"][:2],
                      "rows": 128})  # first two block-rows
    ref = (A @ B1.T.astype(np.float64)).T
    assert np.array_equal(outs[0][:, :128], ref.astype(np.float32))


@pytest.mark.parametrize("mode", ["single", "pair"])
@pytest.mark.parametrize("precision,max_chain", [("tf32", 8), ("tf32", 40), ("bf16", 16), ("fp16", 100)])
def test_bounded_chains_bit_exact(oracle, lib, precision, max_chain, mode):
    """max_chain cuts super-rows into passes folded through the master accumulators in TMEM
    (sched_types.h Item); integer operands must still be exact, whatever the cut."""
    rng = np.random.default_rng(31)
    heights = [64, 30, 64, 64, 7, 64, 64, 64, 16, 64, 1, 64]
    v = random_vbr(rng, len(heights), 2048, 64, heights, 0.6, values="int")
    n = 300
    Bm = rng.integers(-3, 4, size=(n, 2048)).astype(np.float32)
    Cg = gpu_multiply(v, Bm, n, precision, max_chain=max_chain, **MODES[mode])
    assert np.array_equal(Cg, oracle.vbr_multiply(v, Bm, n))


def test_tf32_long_positive_sums_within_1e5(lib):
    """The tensor core's fp32 accumulation truncates, so an all-positive sum of thousands of MMAs
    drifts (5.6e-5 measured at BASELINE config #3).  The tf32 default bounds the chain; the result
    must stay within 1e-5 of fp64 on the tf32-rounded operands even for 512 blocks per row."""
    rng = np.random.default_rng(32)
    block_rows, cols, w, n = 8, 32768, 64, 128
    v = random_vbr(rng, block_rows, cols, w, [64] * block_rows, 1.0, values="ones")
    Bm = rng.random((n, cols), dtype=np.float32)
    ref = round_to(Bm, "tf32").astype(np.float64) @ vbr_to_dense(v).T       # [n, rows]
    Cg = gpu_multiply(v, Bm, n, "tf32")
    assert rel_err(Cg, ref) <= TOL_ROUNDED
    # the unbounded chain is what the limit protects against: it must not be MORE accurate
    Cu = gpu_multiply(v, Bm, n, "tf32", max_chain=-1)
    assert rel_err(Cu, ref) >= rel_err(Cg, ref)
    assert rel_err(Cu, ref) <= TOL_UNROUNDED["tf32"]


WIDE_CASES = [
    # block_rows, cols, w, heights, density, n, wide_tiles, extra
    (7, 128, 64, [64] * 7, 0.4, 600, 2, {}),
    (11, 256, 64, [64, 16, 64, 48, 64, 64, 32, 64, 64, 80, 64], 0.5, 1100, 4, {"row_order": 1}),
    (40, 64, 8, [1] * 40, 0.3, 520, 4, {"gather_max_height": -1}),      # tail tiles beyond n, height-1 block-rows on MMAs
    (48, 2048, 64, [64] * 45 + [63, 30, 7], 0.12, 1500, 2, {"num_ctas": 8}),   # many items per worker, the slots wrap
    (48, 2048, 64, [64] * 48, 0.12, 2048, 4, {}),                       # the ER shape of config #2, scaled down
    (9, 640, 64, [64, 63, 30, 64, 1, 128, 7, 64, 33], 0.7, 777, 2, {"gather_max_height": -1}),
]


@pytest.mark.parametrize("mode", ["single", "pair"])
@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("case", range(len(WIDE_CASES)))
def test_wide_items_bit_exact(oracle, lib, case, precision, mode):
    """wide_tiles = 2 / 4: a work item covers that many column tiles, every pipeline stage carries that many
    panels of B for one set of A images, the tiles accumulate side by side in TMEM (512 / wide_tiles columns
    each) and are drained one after the other.  Integer operands: bit-exact."""
    block_rows, cols, w, heights, density, n, tiles, extra = WIDE_CASES[case]
    rng = np.random.default_rng(900 + case)
    v = random_vbr(rng, block_rows, cols, w, heights, density, values="int")
    Bm = rng.integers(-3, 4, size=(n, cols)).astype(np.float32)
    h = sparta_b200.Handle.from_vbr(v["rows"], cols, w, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                    precision=precision, wide_tiles=tiles, max_chain=-1, **dict(extra, **MODES[mode]))
    try:
        h.set_B(Bm, cols, n)
        assert h.stats()["wide_tiles"] == tiles
        h.run()
        a = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"]).copy()
        h.run()
        b = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"])
    finally:
        h.close()
    Cref = oracle.vbr_multiply(v, Bm, n)
    assert np.array_equal(a, Cref) and np.array_equal(b, Cref)


def test_wide_items_are_chosen_for_er_like_lists_only(oracle, lib):
    """n_hint between 2 and 6 tile widths: block-rows sharing few column blocks (ER, 10 % block density) get
    wide items, dense lists (many members per column block) keep one tile per item; from 6 widths on every
    bf16 / fp16 handle gets two tiles; tf32 chains that need fold passes keep one.  Real operands: within
    the usual tolerances."""
    rng = np.random.default_rng(77)
    er = random_vbr(rng, 64, 4096, 64, [64] * 64, 0.1, values="real")
    dense = random_vbr(rng, 16, 1024, 64, [64] * 16, 0.9, values="real")
    n = 1024
    for v, cols, expect in ((er, 4096, 2), (dense, 1024, 1)):
        Bm = rng.standard_normal((n, cols)).astype(np.float32)
        h = sparta_b200.Handle.from_vbr(v["rows"], cols, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                        precision="bf16", n_hint=n)
        try:
            assert h.stats()["wide_tiles"] == expect
            h.set_B(Bm, cols, n)
            h.run()
            Cg = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"])
        finally:
            h.close()
        Cr = oracle.vbr_multiply(rounded(v, "bf16"), round_to(Bm, "bf16"), n)
        assert rel_err(Cg, Cr) <= TOL_ROUNDED
    h = sparta_b200.Handle.from_vbr(er["rows"], 4096, 64, er["row_part"], er["nzcount"], er["jab"], er["mab"],
                                    precision="bf16", n_hint=256)
    try:
        assert h.stats()["wide_tiles"] == 1       # n spans one tile: nothing to share
    finally:
        h.close()
    for precision, chain, expect in (("bf16", 0, 2), ("tf32", 64, 1)):   # chains cut at 64 MMAs need the master accumulators
        h = sparta_b200.Handle.from_vbr(dense["rows"], 1024, 64, dense["row_part"], dense["nzcount"], dense["jab"],
                                        dense["mab"], precision=precision, n_hint=2048, max_chain=chain)
        try:
            assert h.stats()["wide_tiles"] == expect
        finally:
            h.close()


def test_wide_items_split_accumulate_and_row_major(oracle, lib):
    """Wide items cut into split pieces (red.add into tiles zeroed in-kernel: T x 128-column pieces per CTA),
    with beta = 1 and with a row-major C."""
    rng = np.random.default_rng(43)
    heights = [64] * 6
    v = random_vbr(rng, len(heights), 8192, 64, heights, 0.3, values="int")
    n = 1000
    Bm = rng.integers(-3, 4, size=(n, 8192)).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n)
    for tiles in (2, 4):
        h = sparta_b200.Handle.from_vbr(v["rows"], 8192, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                        precision="bf16", split_k=2, wide_tiles=tiles)
        try:
            h.set_B(Bm, 8192, n)
            st = h.stats()
            assert st["split_pieces"] > 0 and st["zero_tiles"] > 0 and st["wide_tiles"] == tiles
            h.run()
            a = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"]).copy()
            h.run()
            b = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"])
        finally:
            h.close()
        assert np.array_equal(a, Cref) and np.array_equal(b, Cref)


@pytest.mark.parametrize("mode", ["single", "pair"])
@pytest.mark.parametrize("precision,max_chain", [("bf16", 0), ("fp16", 0), ("tf32", 24), ("tf32", 0)])
def test_split_pieces_bit_exact(oracle, lib, precision, max_chain, mode):
    """split_k = 2: the column-block lists are cut into one piece per worker and the pieces add
    their partial sums to C with red.global.add.f32 after zero_c_tiles_kernel (sched_types.h).
    Integer operands: exact whatever the order of the additions.  Launching again must give the
    same C (the tiles are re-zeroed before every launch)."""
    rng = np.random.default_rng(41)
    heights = [64, 64, 64, 30, 64, 64, 64, 64, 64, 17, 64, 64, 1, 64]
    v = random_vbr(rng, len(heights), 4096, 64, heights, 0.7, values="int")
    n = 300
    Bm = rng.integers(-3, 4, size=(n, 4096)).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n)
    h = sparta_b200.Handle.from_vbr(v["rows"], 4096, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                    precision=precision, max_chain=max_chain, split_k=2, **MODES[mode])
    try:
        h.set_B(Bm, 4096, n)
        st = h.stats()
        assert st["split_pieces"] > 0 and st["zero_tiles"] > 0
        h.run()
        a = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"]).copy()
        h.run()
        b = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"])
    finally:
        h.close()
    assert np.array_equal(a, Cref) and np.array_equal(b, Cref)


def test_split_pieces_accumulate_and_row_major(oracle, lib):
    """Split pieces with beta = 1 (no zeroing: the pieces add onto the caller's C) and with the
    Blocked-ELL path's row-major C."""
    rng = np.random.default_rng(42)
    v = random_vbr(rng, 10, 2048, 64, [64] * 10, 0.8, values="int")
    n = 520
    Bm = rng.integers(-3, 4, size=(n, 2048)).astype(np.float32)
    C0 = rng.integers(-5, 6, size=(n, v["rows"])).astype(np.float32)
    h = sparta_b200.Handle.from_vbr(v["rows"], 2048, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                    accumulate=1, split_k=2)
    h.set_B(Bm, 2048, n)
    h.set_C(C0, v["rows"])
    assert h.stats()["split_pieces"] > 0
    h.run()
    out = h.get_C(np.zeros_like(C0), v["rows"])
    h.close()
    assert np.array_equal(out, oracle.vbr_multiply(v, Bm, n, C_init=C0))
    # row-major B and C (b_layout = c_layout = 2)
    h = sparta_b200.Handle.from_vbr(v["rows"], 2048, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                    split_k=2, b_layout=2, c_layout=2)
    B_rm = np.ascontiguousarray(Bm.T)
    h.set_B(B_rm, n, n)
    assert h.stats()["split_pieces"] > 0
    h.run()
    out = h.get_C(np.zeros((v["rows"], n), np.float32), n)
    h.close()
    assert np.array_equal(out, oracle.vbr_multiply(v, Bm, n).T)


def test_split_real_operands_within_tolerance(oracle, lib):
    rng = np.random.default_rng(43)
    v = random_vbr(rng, 12, 4096, 64, [64] * 12, 0.9, values="uniform")
    n = 256
    Bm = rng.random((n, 4096), dtype=np.float32)
    for precision in ("bf16", "tf32"):
        Cg = gpu_multiply(v, Bm, n, precision, split_k=2)
        C_rounded = oracle.vbr_multiply(rounded(v, precision), round_to(Bm, precision), n)
        assert rel_err(Cg, C_rounded) <= TOL_ROUNDED
        assert rel_err(Cg, oracle.vbr_multiply(v, Bm, n)) <= TOL_UNROUNDED[precision]


@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("layout", ["col", "row"])
def test_short_block_rows_on_the_gather_kernel(oracle, lib, precision, layout):
    """Block-rows of at most gather_max_height rows (default 7) leave the tile schedule and run, as
    the nonzeros of their blocks, on the gather kernel of the family; the rest stays on the tensor
    cores.  Same product: integer operands bit-exact with gathering on, off and at another
    threshold; real operands within the stated tolerances; accumulate = 1; both C layouts."""
    rng = np.random.default_rng(77)
    heights = [1, 1, 64, 3, 1, 7, 8, 1, 1, 30, 2, 1, 1, 1, 5, 64, 1, 4, 1, 6, 1, 1, 16, 1] * 3
    v = random_vbr(rng, len(heights), 1000, 64, heights, 0.35, values="int")     # cols % w != 0
    # make the short blocks sparse inside, like the blocks of a clustered sparse matrix
    keep = rng.random(v["mab"].shape) < 0.1
    v["mab"] = (v["mab"] * keep).astype(np.float32)
    n = 200
    Bm = rng.integers(-3, 4, size=(n, 1000)).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n)
    lay = dict(b_layout=2, c_layout=2) if layout == "row" else {}

    def run(values_v, B, accumulate=0, C0=None, **opts):
        h = sparta_b200.Handle.from_vbr(v["rows"], 1000, 64, values_v["row_part"], values_v["nzcount"], values_v["jab"],
                                        values_v["mab"], precision=precision, accumulate=accumulate, **lay, **opts)
        try:
            if layout == "row":
                h.set_B(np.ascontiguousarray(B.T), n, n)
            else:
                h.set_B(B, 1000, n)
            if C0 is not None:
                h.set_C(np.ascontiguousarray(C0.T) if layout == "row" else C0, n if layout == "row" else v["rows"])
            h.run()
            st = h.stats()
            if layout == "row":
                out = h.get_C(np.zeros((v["rows"], n), np.float32), n).T.copy()
            else:
                out = h.get_C(np.zeros((n, v["rows"]), np.float32), v["rows"])
            return out, st
        finally:
            h.close()

    on, st_on = run(v, Bm)
    off, st_off = run(v, Bm, gather_max_height=-1)
    upto16, st16 = run(v, Bm, gather_max_height=16)
    # the gather rows walked in 5 ranges of k (one launch each; pass 0 writes, the others add)
    passes5, st5 = run(v, Bm, gather_passes=5)
    # (ranges are multiples of 64 columns: 1000 columns make 4 non-empty ranges of 256)
    assert np.array_equal(passes5, Cref) and st5["kernel_launches"] == 1 + 4
    hs = np.asarray(heights)
    assert st_on["gather_rows"] == int(hs[hs <= 7].sum()) and st_off["gather_rows"] == 0
    assert st16["gather_rows"] == int(hs[hs <= 16].sum())
    assert st_on["gather_nnz"] > 0 and st_on["nztot"] == st_off["nztot"] == v["mab"].size
    assert st_on["nz_blocks"] == st_off["nz_blocks"] == len(v["jab"])
    assert st_on["chunks"] < st_off["chunks"]          # the short block-rows left the tile schedule
    for out in (on, off, upto16):
        assert np.array_equal(out, Cref)
    # beta = 1
    C0 = rng.integers(-5, 6, size=(n, v["rows"])).astype(np.float32)
    acc, _ = run(v, Bm, accumulate=1, C0=C0)
    assert np.array_equal(acc, Cref + C0)
    # real values: both kernel families round their operands to the handle's precision
    vr = dict(v)
    # same sparsity as the integer matrix (its mask also zeroes the columns past the matrix edge in the
    # last column block, as the reference's fill does, vbr.cpp:205-227)
    vr["mab"] = (rng.uniform(-1, 1, size=v["mab"].shape) * (v["mab"] != 0)).astype(np.float32)
    Br = rng.uniform(0, 1, size=(n, 1000)).astype(np.float32)
    got, _ = run(vr, Br)
    assert rel_err(got, oracle.vbr_multiply(rounded(vr, precision), round_to(Br, precision), n)) <= TOL_ROUNDED
    assert rel_err(got, oracle.vbr_multiply(vr, Br, n)) <= TOL_UNROUNDED[precision]


BA_CASES = [
    (6, 96, 16, [16] * 6, 0.5, 40),
    (5, 70, 16, [3, 17, 1, 64, 30], 0.6, 130),      # ragged heights = ragged k extents; cols % w != 0
    (3, 300, 100, [64, 64, 20], 0.8, 16),
    (3, 64, 32, [200, 70, 9], 1.0, 8),
    (24, 1024, 64, [64] * 24, 0.5, 384),
    (40, 2048, 64, [64] * 37 + [63, 30, 7], 0.5, 700),
]


@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("shape", [(12, 64, 64), (9, 16, 16), (1, 64, 23)])
def test_inverted_product_equals_the_reference_loop(oracle, lib, precision, shape):
    """The kernel against the reference's OWN loop for -M 6 (cublas_blockmat_multiplyBA,
    src/cuda/cuda_utilities.cpp:640-690, restated index for index in the oracle) in the regime where
    that loop's B offset is well defined: B_rows = 1 with constant heights h == w (-b == -B), and a
    single block-row of any height (tests/test_oracle.py shows the two products coincide exactly
    there and why they differ elsewhere).  Integer operands: exact."""
    block_rows, w, hgt = shape
    rng = np.random.default_rng(900 + block_rows)
    cols = w * 11
    v = random_vbr(rng, block_rows, cols, w, [hgt] * block_rows, 0.6, values="int", empty_rows=False)
    m = 1 if block_rows > 1 else 37
    Bt = rng.integers(-3, 4, size=(v["rows"], m)).astype(np.float32)
    lit, out_of_range = oracle.ref_multiplyBA_literal(v, Bt, m)
    assert not out_of_range
    h = sparta_b200.Handle.from_vbr_BA(v["rows"], cols, w, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                       precision=precision)
    try:
        h.set_B(Bt, m, m)
        h.run()
        out = h.get_C(np.zeros((cols, m), np.float32), m)
    finally:
        h.close()
    assert np.array_equal(out, lit[:cols])


@pytest.mark.parametrize("mode", ["single", "pair"])
@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("case", range(len(BA_CASES)))
def test_inverted_product_bit_exact(oracle, lib, case, precision, mode):
    """C = B*A (-M 6 / -M 11 replacement, sparta_vbr_create_BA): integer operands, exact."""
    block_rows, cols, w, heights, density, m = BA_CASES[case]
    rng = np.random.default_rng(400 + case)
    v = random_vbr(rng, block_rows, cols, w, heights, density, values="int")
    Bt = rng.integers(-3, 4, size=(v["rows"], m)).astype(np.float32)      # row k = column k of B
    h = sparta_b200.Handle.from_vbr_BA(v["rows"], cols, w, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                       precision=precision, **MODES[mode])
    try:
        h.set_B(Bt, m, m)
        st = h.stats()
        assert st["rows"] == cols and st["nztot"] == v["mab"].size
        h.run()
        out = h.get_C(np.zeros((cols, m), np.float32), m)
    finally:
        h.close()
    assert np.array_equal(out, oracle.vbr_multiply_BA(v, Bt, m))


def test_inverted_product_one_shot_and_tolerance(oracle, lib):
    from sparta_b200.api import vbr_spmm_BA
    rng = np.random.default_rng(44)
    heights = [64] * 9 + [30]
    v = random_vbr(rng, len(heights), 640, 64, heights, 0.6, values="uniform")
    m = 200
    Bt = rng.random((v["rows"], m), dtype=np.float32)
    A = VBR(v["rows"], v["cols"], 64, v["row_part"], v["nzcount"], v["jab"], v["mab"])
    for precision in ("bf16", "tf32"):
        Cg, dt = vbr_spmm_BA(A, Bt, m, precision)
        assert dt > 0
        C_rounded = oracle.vbr_multiply_BA(rounded(v, precision), round_to(Bt, precision), m)
        assert rel_err(Cg, C_rounded) <= TOL_ROUNDED
        assert rel_err(Cg, oracle.vbr_multiply_BA(v, Bt, m)) <= TOL_UNROUNDED[precision]


def test_inverted_product_column_block_shards(oracle, lib):
    """Shards of the inverted product are ranges of column blocks: slabs of C's columns."""
    rng = np.random.default_rng(45)
    v = random_vbr(rng, 12, 512, 32, [32] * 12, 0.5, values="int")
    m = 64
    Bt = rng.integers(-2, 3, size=(v["rows"], m)).astype(np.float32)
    Cref = oracle.vbr_multiply_BA(v, Bt, m)
    slabs = []
    for lo, hi in ((0, 5), (5, 11), (11, 16)):
        h = sparta_b200.Handle.from_vbr_BA(v["rows"], 512, 32, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                           block_row_begin=lo, block_row_end=hi)
        h.set_B(Bt, m, m)
        h.run()
        slabs.append(h.get_C(np.zeros(((hi - lo) * 32, m), np.float32), m))
        h.close()
    assert np.array_equal(np.concatenate(slabs, axis=0), Cref)


@pytest.mark.parametrize("layout", ["col_major", "row_major"])
def test_unpermuted_read_back_equals_csr_multiply(oracle, lib, layout):
    """sparta_get_C_permuted with the reference's get_permutation puts C back in the original row
    order: on a square matrix that is exactly CSR::multiply's result (src/general/csr.cpp:49-65),
    which never reorders anything."""
    from sparta_b200.lib import host_permutation
    res = oracle.run(os.path.join(GOLDEN, "rmat8.el"), P=1, a=5, b=16, B=16, t=0.6)
    rows = res["rows"]
    assert rows == res["cols"]
    perm = host_permutation(res["grouping"])
    assert np.array_equal(perm, oracle.permutation(res["grouping"]))
    rng = np.random.default_rng(51)
    n = 72
    Bm = rng.integers(-3, 4, size=(n, rows)).astype(np.float32)              # [n, cols]: column-major B
    Cref = oracle.csr_multiply(rows, res["csr_rowptr"], res["csr_colind"], res["csr_val"], True, Bm, n)
    opts = {} if layout == "col_major" else dict(b_layout=2, c_layout=2)
    h = sparta_b200.Handle.from_vbr(rows, res["cols"], 16, res["row_part"], res["nzcount"], res["jab"], res["mab"], **opts)
    try:
        if layout == "col_major":
            h.set_B(Bm, rows, n)
            h.run()
            blocked = h.get_C(np.zeros((n, rows), np.float32), rows).copy()
            out = h.get_C_permuted(np.full((n, rows), -1.0, np.float32), rows, perm, rows)
            assert np.array_equal(out, Cref)
            assert np.array_equal(blocked, oracle.vbr_multiply(res, Bm, n))   # get_C stays in blocked order
        else:
            h.set_B(np.ascontiguousarray(Bm.T), n, n)
            h.run()
            out = h.get_C_permuted(np.full((rows, n), -1.0, np.float32), n, perm, rows)
            assert np.array_equal(out, Cref.T)
    finally:
        h.close()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_variable_height_blocking_end_to_end(oracle, lib, precision):
    """BASELINE config #4's shape at small scale: edge list -> the library's own `-a 4` clustering
    (bit-exact with the reference, tests/test_blocking.py) -> variable-height VBR (most block-rows
    one row tall, a few hundreds of rows tall) -> multiply, against the oracle on the same VBR."""
    from sparta_b200.lib import host_blocking, host_vbr_fill
    N = 1 << 12
    r, c = synth.rmat_edges(12, int(4e-3 * N * N), seed=3)
    r, c = synth.pin_shape(r, c, N, N)
    rowptr, colind, _ = synth.csr_from_edges(r, c, N)
    g = host_blocking(N, N, rowptr, colind, algo=4, tau=0.6, block_col_size=64, row_block_size=64,
                      sim_measure=1, use_pattern=True, use_group=False)
    v = host_vbr_fill(N, N, rowptr, colind, None, g, 64, 64, force_fixed_size=False, pattern_only=True)
    heights = np.diff(v["row_part"])
    assert (heights == 1).mean() > 0.25 and heights.max() > 64         # the shape the test is about
    n = 320
    rng = np.random.default_rng(61)
    Bm = rng.integers(-3, 4, size=(n, N)).astype(np.float32)
    Cg = gpu_multiply(v, Bm, n, precision)
    assert np.array_equal(Cg, oracle.vbr_multiply(v, Bm, n))
    # real operands, tf32 tolerance of the north star (<= 1e-5 against the rounded operands)
    Br = rng.random((n, N), dtype=np.float32)
    Cg = gpu_multiply(v, Br, n, precision)
    assert rel_err(Cg, oracle.vbr_multiply(v, round_to(Br, precision), n)) <= TOL_ROUNDED


@pytest.mark.parametrize("mode", ["single", "pair"])
@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_fused_short_block_rows_bit_exact(oracle, lib, precision, mode):
    """fuse_rows on (default) and off give the oracle's product on a VBR full of short block-rows."""
    rng = np.random.default_rng(82)
    heights = [1, 1, 2, 1, 5, 1, 1, 1, 1, 3, 64, 1, 1, 1, 9, 8, 1, 1, 40] + [1] * 60 + [2, 2, 3, 130, 1, 1]
    v = random_vbr(rng, len(heights), 1024, 64, heights, 0.3, values="int")
    n = 264
    Bm = rng.integers(-3, 4, size=(n, 1024)).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n)
    assert np.array_equal(gpu_multiply(v, Bm, n, precision, **MODES[mode]), Cref)
    assert np.array_equal(gpu_multiply(v, Bm, n, precision, fuse_rows=1, **MODES[mode]), Cref)


def test_sparse_upload_matches_dense_upload(oracle, lib, monkeypatch):
    """Sources above 4 M elements whose sample is < 10 % nonzero cross PCIe as (offset, value)
    pairs and are rebuilt on the device (abi.cu, scan_nonzeros); the product must be the one the
    dense upload gives, including negative zeros and denormal-free edge values."""
    rng = np.random.default_rng(91)
    heights = [64] * 40 + [30, 1, 1, 7]
    v = random_vbr(rng, len(heights), 4096, 64, heights, 0.6, values="int")
    assert v["mab"].size > (1 << 22)
    keep = rng.random(v["mab"].size) < 0.03
    v["mab"] = np.where(keep, v["mab"], 0.0).astype(np.float32)
    v["mab"][::1001] = -0.0                     # negative zeros are zeros
    n = 136
    Bm = rng.integers(-3, 4, size=(n, 4096)).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n)
    sparse = gpu_multiply(v, Bm, n, "bf16")
    monkeypatch.setenv("SPARTA_DENSE_UPLOAD", "1")
    dense = gpu_multiply(v, Bm, n, "bf16")
    assert np.array_equal(sparse, Cref) and np.array_equal(dense, Cref)


def test_shards_scatter_into_one_matrix_in_original_order(oracle, lib):
    """Multi-GPU read-back on one device: every shard scatters its slab of C straight into a
    full-size DEVICE matrix in the original row order (sparta_get_C_permuted, on_device = 1)."""
    import ctypes as C
    import torch
    from sparta_b200 import lib as L
    from sparta_b200.lib import host_permutation
    res = oracle.run(os.path.join(GOLDEN, "rmat8.el"), P=1, a=5, b=16, B=16, t=0.6)
    rows = res["rows"]
    perm = host_permutation(res["grouping"])
    rng = np.random.default_rng(52)
    n = 40
    Bm = rng.integers(-3, 4, size=(n, rows)).astype(np.float32)
    Cref = oracle.csr_multiply(rows, res["csr_rowptr"], res["csr_colind"], res["csr_val"], True, Bm, n)
    full = torch.full((n, rows), -1.0, dtype=torch.float32, device="cuda")
    cuts = sparta_b200.partition_block_rows_modelled(rows, res["cols"], 16, res["row_part"], res["nzcount"], res["jab"], n, 3)
    for i in range(3):
        lo, hi = int(cuts[i]), int(cuts[i + 1])
        h = sparta_b200.Handle.from_vbr(rows, res["cols"], 16, res["row_part"], res["nzcount"], res["jab"], res["mab"],
                                        block_row_begin=lo, block_row_end=hi)
        h.set_B(Bm, rows, n)
        h.run()
        r0, r1 = int(res["row_part"][lo]), int(res["row_part"][hi])
        row_map = np.ascontiguousarray(perm[r0:r1], dtype=np.int64)
        if r1 > r0:
            L._check(sparta_b200.load().sparta_get_C_permuted(h._h, C.c_void_p(full.data_ptr()), rows, L._ptr(row_map),
                                                              rows, 1))
        h.close()
    torch.cuda.synchronize()
    assert np.array_equal(full.cpu().numpy(), Cref)


@pytest.mark.parametrize("flags", [dict(a=5, b=16, B=16, t=0.6), dict(a=4, b=8, B=8, t=0.5), dict(a=3, b=16, B=16, t=0.4, F=1),
                                   dict(a=2, b=64, B=64, F=1)],
                         ids=lambda f: "-".join(f"{k}{v}" for k, v in f.items()))
@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_handle_from_csr_and_grouping(oracle, lib, flags, precision):
    """sparta_vbr_create_from_csr: A from the flat CSR + the row grouping, dense blocks rebuilt on the
    device (the fp32 mab never exists on the host).  Against (i) the handle built from the VBR arrays
    of sparta_host_vbr_fill -- same C bit for bit, same nztot / block counts -- and (ii) the oracle's
    VBR::multiply; weighted values, -F 1 padding, variable heights (gather rows), shards."""
    from sparta_b200 import lib as L
    rng = np.random.default_rng(5)
    r, c = synth.rmat_edges(10, 9000, seed=4)
    keep = c < 1000
    r, c = synth.pin_shape(r[keep], c[keep], 1024, 1000)
    vals = rng.integers(-3, 4, size=len(r)).astype(np.float32)
    vals[vals == 0] = 1
    rowptr, colind, val = synth.csr_from_edges(r, c, 1024, vals)
    f = dict(a=3, b=3, B=3, t=0.1, F=0)
    f.update(flags)
    g = L.host_blocking(1024, 1000, rowptr, colind, algo=f["a"], tau=f["t"], block_col_size=f["b"],
                        row_block_size=f["B"], sim_measure=1, use_pattern=True, force_fixed_size=bool(f["F"]))
    v = L.host_vbr_fill(1024, 1000, rowptr, colind, val, g, f["b"], f["B"], bool(f["F"]))
    n = 96
    Bm = rng.integers(-3, 4, size=(n, v["cols"])).astype(np.float32)
    Cref = oracle.vbr_multiply(v, Bm, n)

    def product(h, rows):
        try:
            h.set_B(Bm, v["cols"], n)
            h.run()
            return h.get_C(np.zeros((n, rows), np.float32), rows), h.stats()
        finally:
            h.close()

    a, st_a = product(sparta_b200.Handle.from_vbr(v["rows"], v["cols"], f["b"], v["row_part"], v["nzcount"], v["jab"],
                                                  v["mab"], precision=precision), v["rows"])
    hc = sparta_b200.Handle.from_csr_grouping(1024, 1000, rowptr, colind, val, g, f["b"], f["B"], bool(f["F"]),
                                              precision=precision)
    assert hc.vbr_dims.tolist() == [v["rows"], v["cols"], v["block_rows"], v["block_cols"], f["b"], v["nztot"]]
    b, st_b = product(hc, v["rows"])
    assert np.array_equal(a, Cref) and np.array_equal(b, Cref)
    for k in ("rows", "block_rows", "nz_blocks", "nztot", "chunks", "items", "gather_rows", "gather_nnz"):
        assert st_a[k] == st_b[k], k
    # pattern-only (val = NULL) and two shards
    vp = L.host_vbr_fill(1024, 1000, rowptr, colind, None, g, f["b"], f["B"], bool(f["F"]), pattern_only=True)
    Cp = oracle.vbr_multiply(vp, Bm, n)
    cut = v["block_rows"] // 2
    parts = []
    for lo, hi in ((0, cut), (cut, v["block_rows"])):
        rows_s = int(v["row_part"][hi] - v["row_part"][lo])
        hs = sparta_b200.Handle.from_csr_grouping(1024, 1000, rowptr, colind, None, g, f["b"], f["B"], bool(f["F"]),
                                                  precision=precision, block_row_begin=lo, block_row_end=hi)
        if rows_s == 0:
            hs.close()
            continue
        parts.append(product(hs, rows_s)[0])
    assert np.array_equal(np.concatenate(parts, axis=1), Cp)


def test_one_shot_from_csr(oracle, lib):
    """sparta_csr_vbr_spmm: host CSR + grouping + B in, host C out, real values within tolerance."""
    import ctypes as C
    from sparta_b200 import lib as L
    rng = np.random.default_rng(6)
    r, c = synth.rmat_edges(10, 12000, seed=8)
    r, c = synth.pin_shape(r, c, 1024, 1024)
    rowptr, colind, val = synth.csr_from_edges(r, c, 1024, rng.uniform(-1, 1, size=len(r)))
    g = L.host_blocking(1024, 1024, rowptr, colind, algo=5, tau=0.6, block_col_size=32, row_block_size=32, sim_measure=1,
                        use_pattern=True)
    v = L.host_vbr_fill(1024, 1024, rowptr, colind, val, g, 32, 32, False)
    n = 130
    Bm = rng.uniform(0, 1, size=(n, 1024)).astype(np.float32)
    out = np.zeros((n, 1024), np.float32)
    dt = C.c_float(0)
    g64 = np.ascontiguousarray(g, dtype=np.int64)
    L._check(lib.sparta_csr_vbr_spmm(1024, 1024, L._ptr(rowptr), L._ptr(colind), L._ptr(val), L._ptr(g64), 32, 32, 0,
                                     L._ptr(Bm), 1024, n, L._ptr(out), 1024, L.PRECISIONS["bf16"], C.byref(dt)))
    assert dt.value > 0
    assert rel_err(out, oracle.vbr_multiply(rounded(v, "bf16"), round_to(Bm, "bf16"), n)) <= TOL_ROUNDED
    assert rel_err(out, oracle.vbr_multiply(v, Bm, n)) <= TOL_UNROUNDED["bf16"]


@pytest.mark.parametrize("gather_c", [0, 1])
def test_multi_gpu_one_shot(oracle, lib, gather_c):
    """sparta_vbr_spmm_multi on 2 GPUs of the box (skipped on a single-GPU box): block-row shards, B
    uploaded once + one NCCL broadcast, C slabs straight into the caller's matrix or all-gathered."""
    import ctypes as C
    from sparta_b200 import lib as L
    if lib.sparta_device_count() < 2:
        pytest.skip("needs two sm_100 devices")
    rng = np.random.default_rng(12)
    heights = [64] * 20 + [30, 1, 1, 7, 64, 64, 5, 1]
    v = random_vbr(rng, len(heights), 2048, 64, heights, 0.5, values="int")
    n = 320
    Bm = rng.integers(-3, 4, size=(n, 2048)).astype(np.float32)
    out = np.zeros((n, v["rows"]), np.float32)
    dt, bc = C.c_float(0), C.c_float(0)
    rp, nz, jab, mab = L._i64(v["row_part"]), L._i64(v["nzcount"]), L._i64(v["jab"]), L._f32(v["mab"])
    L._check(lib.sparta_vbr_spmm_multi(v["rows"], 2048, len(heights), 64, L._ptr(rp), L._ptr(nz), L._ptr(jab), L._ptr(mab),
                                       L._ptr(Bm), 2048, n, L._ptr(out), v["rows"], L.PRECISIONS["bf16"], 2, gather_c,
                                       C.byref(dt), C.byref(bc)))
    assert np.array_equal(out, oracle.vbr_multiply(v, Bm, n))
    assert dt.value > 0 and bc.value > 0
