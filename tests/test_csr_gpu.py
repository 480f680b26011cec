"""Parity of the sm_100a CSR x dense kernel (`-M 2` replacement, csrc/csr_kernel.cu), called
through the C ABI, against the oracle's restatement of CSR::multiply (src/general/csr.cpp:49-65).

  * precision "tf32" on this path = plain fp32 multiply-then-add in the reference's own order:
    BIT-IDENTICAL to CSR::multiply for any operands;
  * bf16 / fp16: B is rounded to 2 bytes, A stays fp32, fp32 accumulation: integer-valued operands
    bit-exact; real operands <= 1e-5 against fp64 on the rounded B and <= 2e-2 against the oracle
    fed the same fp32 inputs (norm max|dC| / max|C|).

CSR::multiply indexes B with `rows` as leading dimension (csr.cpp:61), so it is only defined for
square A; rectangular cases are checked against fp64 (exact for the integer operands used).
"""
import os

import numpy as np
import pytest

import sparta_b200
from sparta_b200 import synth
from sparta_b200.api import csr_spmm
from tests.util import rel_err, round_to

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def random_csr(rng, rows, cols, density, values="int", heavy_rows=()):
    """Flat CSR with ascending columns per row; `heavy_rows` get every column (long rows)."""
    rowptr = [0]
    colind, val = [], []
    for i in range(rows):
        if i in heavy_rows:
            cs = np.arange(cols)
        else:
            cs = np.nonzero(rng.random(cols) < density)[0]
        colind.append(cs)
        if values == "int":
            val.append(rng.integers(-2, 3, size=len(cs)).astype(np.float32))
        else:
            val.append(rng.uniform(-1, 1, size=len(cs)).astype(np.float32))
        rowptr.append(rowptr[-1] + len(cs))
    return (np.asarray(rowptr, np.int64), np.concatenate(colind).astype(np.int64) if rowptr[-1] else np.zeros(0, np.int64),
            np.concatenate(val).astype(np.float32) if rowptr[-1] else np.zeros(0, np.float32))


def dense_of(rows, cols, rowptr, colind, val):
    A = np.zeros((rows, cols), dtype=np.float64)
    for i in range(rows):
        A[i, colind[rowptr[i]:rowptr[i + 1]]] = val[rowptr[i]:rowptr[i + 1]]
    return A


# rows, cols, density, n, heavy rows
CASES = [
    (9, 9, 0.3, 2, ()),
    (64, 64, 0.2, 8, (3,)),
    (100, 100, 0.1, 33, (0, 99)),          # n not a multiple of 8: predicated tail stores
    (257, 257, 0.05, 300, (128,)),         # two column tiles, rows not a multiple of 8
    (40, 300, 0.3, 64, (5,)),              # rectangular, rows of > 64 entries
    (300, 40, 0.3, 520, ()),               # rectangular, three column tiles
    (16, 16, 0.0, 16, ()),                 # empty matrix -> C = 0
    (33, 33, 0.5, 1, ()),                  # a single column of B
]


@pytest.mark.parametrize("precision", ["bf16", "fp16", "tf32"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_integer_operands_bit_exact(oracle, lib, case, precision):
    rows, cols, density, n, heavy = CASES[case]
    rng = np.random.default_rng(300 + case)
    rowptr, colind, val = random_csr(rng, rows, cols, density, "int", heavy)
    B = rng.integers(-3, 4, size=(cols, n)).astype(np.float32)       # row-major cols x n
    Cg, dt = csr_spmm(rows, cols, rowptr, colind, val, B, n, precision)
    Cref = (dense_of(rows, cols, rowptr, colind, val) @ B.astype(np.float64)).astype(np.float32)
    assert np.array_equal(Cg, Cref)
    if rows == cols:   # the reference routine itself (column-major B and C, ld = rows)
        Co = oracle.csr_multiply(rows, rowptr, colind, val, False, np.ascontiguousarray(B.T), n)
        assert np.array_equal(Cg, Co.T)


@pytest.mark.parametrize("case", [1, 2, 3])
def test_fp32_mode_bit_identical_to_reference_routine(oracle, lib, case):
    rows, cols, density, n, heavy = CASES[case]
    rng = np.random.default_rng(400 + case)
    rowptr, colind, val = random_csr(rng, rows, cols, density, "uniform", heavy)
    B = rng.random((cols, n), dtype=np.float32)
    Cg, _ = csr_spmm(rows, cols, rowptr, colind, val, B, n, "tf32")
    Co = oracle.csr_multiply(rows, rowptr, colind, val, False, np.ascontiguousarray(B.T), n)
    assert np.array_equal(Cg, Co.T)
    # pattern-only matrices multiply with 1 (csr.cpp:59)
    Cg, _ = csr_spmm(rows, cols, rowptr, colind, None, B, n, "tf32")
    Co = oracle.csr_multiply(rows, rowptr, colind, val, True, np.ascontiguousarray(B.T), n)
    assert np.array_equal(Cg, Co.T)


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("case", [1, 3, 4])
def test_real_operands_within_tolerance(oracle, lib, case, precision):
    rows, cols, density, n, heavy = CASES[case]
    rng = np.random.default_rng(500 + case)
    rowptr, colind, val = random_csr(rng, rows, cols, density, "uniform", heavy)
    B = rng.random((cols, n), dtype=np.float32)
    Cg, _ = csr_spmm(rows, cols, rowptr, colind, val, B, n, precision)
    A = dense_of(rows, cols, rowptr, colind, val)
    assert rel_err(Cg, A @ round_to(B, precision).astype(np.float64)) <= 1e-5
    assert rel_err(Cg, A @ B.astype(np.float64)) <= 2e-2
    if rows == cols:
        Co = oracle.csr_multiply(rows, rowptr, colind, val, False, np.ascontiguousarray(B.T), n)
        assert rel_err(Cg, Co.T) <= 2e-2


@pytest.mark.parametrize("name,pattern", [("rmat8.el", 1), ("er_weighted.el", 0), ("TEST_matrix_weighted.el", 0)])
def test_golden_matrices_against_reference_routine(oracle, lib, name, pattern):
    """The committed test matrices through the reference's reader (which drops line 1,
    csr.cpp:213), multiplied by the kernel and by the restated CSR::multiply."""
    res = oracle.run(os.path.join(GOLDEN, name), P=pattern, a=2, b=16, B=16, fill=False)
    rowptr, colind, val = res["csr_rowptr"], res["csr_colind"], res["csr_val"]
    rows, cols = res["csr_rows"], res["csr_cols"]
    rng = np.random.default_rng(11)
    n = 72
    B = rng.random((cols, n), dtype=np.float32)
    Cg, _ = csr_spmm(rows, cols, rowptr, colind, None if pattern else val, B, n, "tf32")
    if rows == cols:
        Cref = oracle.csr_multiply(rows, rowptr, colind, val, bool(pattern), np.ascontiguousarray(B.T), n).T
        assert np.array_equal(Cg, Cref)
    else:   # CSR::multiply is undefined for rectangular A (csr.cpp:61): fp64 instead
        Cref = dense_of(rows, cols, rowptr, colind, np.ones_like(val) if pattern else val) @ B.astype(np.float64)
        assert rel_err(Cg, Cref) <= 1e-6
    Cg, _ = csr_spmm(rows, cols, rowptr, colind, None if pattern else val, B, n, "bf16")
    assert rel_err(Cg, Cref) <= 2e-2


def test_handle_api_layouts_accumulate_and_row_shards(oracle, lib):
    rng = np.random.default_rng(12)
    rows = cols = 200
    rowptr, colind, val = random_csr(rng, rows, cols, 0.1, "int", (7,))
    n = 40
    A = dense_of(rows, cols, rowptr, colind, val)
    B = rng.integers(-3, 4, size=(cols, n)).astype(np.float32)
    Cref = (A @ B.astype(np.float64)).astype(np.float32)
    # column-major B and C (the layouts of the CPU routine), padded leading dimensions
    ldb, ldc = cols + 8, rows + 4
    Bcm = np.zeros((n, ldb), np.float32)
    Bcm[:, :cols] = B.T
    h = sparta_b200.Handle.from_csr(rows, cols, rowptr, colind, val, precision="tf32",
                                    b_layout=sparta_b200.COL_MAJOR, c_layout=sparta_b200.COL_MAJOR)
    h.set_B(Bcm, ldb, n)
    h.run()
    out = np.full((n, ldc), -7.0, np.float32)
    h.get_C(out, ldc)
    assert np.array_equal(out[:, :rows], Cref.T) and np.all(out[:, rows:] == -7.0)
    h.run(); h.run()
    assert np.array_equal(h.get_C(np.zeros((n, ldc), np.float32), ldc)[:, :rows], Cref.T)
    st = h.stats()
    assert st["kernel_launches"] == 3 and st["nztot"] == len(colind) and st["rows"] == rows
    h.close()
    # accumulate = 1: C += A*B on an uploaded C
    C0 = rng.integers(-5, 6, size=(rows, n)).astype(np.float32)
    h = sparta_b200.Handle.from_csr(rows, cols, rowptr, colind, val, accumulate=1)
    h.set_B(B, n, n)
    h.set_C(C0, n)
    h.run()
    assert np.array_equal(h.get_C(np.zeros((rows, n), np.float32), n), Cref + C0)
    h.close()
    # row shards (the multi-GPU partition on one device) reassemble to the full product
    slabs = []
    for lo, hi in ((0, 0), (0, 50), (50, 51), (51, 51), (51, 200)):
        h = sparta_b200.Handle.from_csr(rows, cols, rowptr, colind, val, block_row_begin=lo, block_row_end=hi)
        if hi == lo:
            assert h.stats()["rows"] == 0
            h.close()
            continue
        h.set_B(B, n, n)
        h.run()
        slabs.append(h.get_C(np.zeros((hi - lo, n), np.float32), n))
        h.close()
    assert np.array_equal(np.concatenate(slabs, axis=0), Cref)


def test_config3_shape_rmat_csr(lib):
    """The R-MAT 65536^2 matrix of BASELINE config #3 as CSR, n = 2048: linearity on integer
    operands (exact) and fp64 recomputation of sampled rows."""
    scale, n = 16, 2048
    N = 1 << scale
    r, c = synth.rmat_edges(scale, int(1e-3 * N * N), seed=1)
    r, c = synth.pin_shape(r, c, N, N)
    rowptr, colind, _ = synth.csr_from_edges(r, c, N)
    rng = np.random.default_rng(3)
    B1 = rng.integers(-3, 4, size=(N, n)).astype(np.float32)
    B2 = rng.integers(-3, 4, size=(N, n)).astype(np.float32)
    h = sparta_b200.Handle.from_csr(N, N, rowptr, colind, None, precision="bf16")
    outs = []
    for Bm in (B1, B2, B1 + B2):
        h.set_B(Bm, n, n)
        h.run()
        outs.append(h.get_C(np.zeros((N, n), np.float32), n).copy())
    h.close()
    assert np.array_equal(outs[0] + outs[1], outs[2])
    nnz_row = np.diff(rowptr)
    for i in [0, int(np.argmax(nnz_row)), N - 1, 12345]:
        ref = B1[colind[rowptr[i]:rowptr[i + 1]]].astype(np.float64).sum(axis=0)
        assert np.array_equal(outs[0][i], ref.astype(np.float32))


def test_fp32_mode_long_rows_real_values(oracle, lib):
    """The fp32 ("tf32") mode is bit-identical to CSR::multiply (src/general/csr.cpp:49-65) for rows of up
    to 512 nonzeros; a longer row is summed as 8 slices whose partial sums are added in order -- the same
    products in a different fp32 association.  Real values: the short rows equal the reference bit for
    bit, the long rows within a few ulp of the row's magnitude."""
    rng = np.random.default_rng(77)
    rows = cols = 1200
    rowptr, colind, val = random_csr(rng, rows, cols, 0.02, "uniform", heavy_rows=(3, 700))
    n = 40
    B = rng.random((cols, n), dtype=np.float32)
    Cg, _ = csr_spmm(rows, cols, rowptr, colind, val, B, n, "tf32")
    Cref = oracle.csr_multiply(rows, rowptr, colind, val, False, np.ascontiguousarray(B.T), n).T
    lens = np.diff(rowptr)
    short = lens <= 512
    assert (~short).sum() == 2 and lens.max() == cols
    assert np.array_equal(Cg[short], Cref[short])
    scale = np.abs(Cref[~short]).max()
    assert np.abs(Cg[~short] - Cref[~short]).max() <= 64 * np.finfo(np.float32).eps * scale
