"""Host-side mirror of the reference's multiply entry points (include/cuda_utilities.h:38-66).

The reference functions take a host `VBR` struct plus host B / C and return the elapsed
milliseconds through `float& dt`.  The functions here keep that data flow (host in, host out,
`dt` = CUDA-event time around the compute only) but go through libsparta_b200's C ABI.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import lib as _lib


@dataclass
class VBR:
    """Field-for-field view of `struct VBR` (reference include/matrices.h:93-122)."""
    rows: int
    cols: int
    block_col_size: int
    row_part: np.ndarray   # int64[block_rows + 1]
    nzcount: np.ndarray    # int64[block_rows]
    jab: np.ndarray        # int64[sum nzcount]
    mab: np.ndarray        # float32[nztot], blocks back to back, column-major inside a block

    @property
    def block_rows(self):
        return len(self.nzcount)

    @property
    def block_cols(self):
        return (self.cols - 1) // self.block_col_size + 1

    @property
    def nztot(self):
        return int(self.mab.size)


def vbr_spmm(A: VBR, B: np.ndarray, B_cols: int, precision="bf16"):
    """C = A*B like cublas_fixed_blocks_multiply / cublas_blockmat_batched
    (reference src/cuda/cuda_utilities.cpp:39, :723): B column-major (ld = A.cols), C
    column-major (ld = A.rows) in blocked row order.  Returns (C as [B_cols, rows] array whose
    row j is column j of C, dt_ms)."""
    lib = _lib.load()
    B = np.ascontiguousarray(B, dtype=np.float32).reshape(-1)
    if B.size < A.cols * B_cols:
        raise ValueError("B must hold cols * B_cols floats (column-major)")
    Cbuf = np.zeros((B_cols, A.rows), dtype=np.float32)
    dt = C.c_float(0)
    rp, nz, jab = (np.ascontiguousarray(x, dtype=np.int64) for x in (A.row_part, A.nzcount, A.jab))
    mab = np.ascontiguousarray(A.mab, dtype=np.float32)
    prec = _lib.PRECISIONS[precision]
    _lib._check(lib.sparta_vbr_spmm(A.rows, A.cols, A.block_rows, A.block_col_size, _lib._ptr(rp),
                                    _lib._ptr(nz), _lib._ptr(jab), _lib._ptr(mab), _lib._ptr(B),
                                    A.cols, B_cols, _lib._ptr(Cbuf), A.rows, prec, C.byref(dt)))
    return Cbuf, dt.value


def vbr_spmm_BA(A: VBR, B: np.ndarray, B_rows: int, precision="bf16"):
    """C = B*A like cublas_blockmat_multiplyBA / cutlas_blockmat_multiplyBA (reference
    src/cuda/cuda_utilities.cpp:553, src/cuda/cutlass_bellpack_lib.cu:542; `-M 6`, `-M 11`):
    B is B_rows x A.rows and C is B_rows x A.cols, both column-major with ld = B_rows; the columns
    of B follow A's blocked row order.  B is passed as the [A.rows, B_rows] array whose row k is
    column k of B; returns (C as [A.cols, B_rows], dt_ms)."""
    lib = _lib.load()
    B = np.ascontiguousarray(B, dtype=np.float32).reshape(-1)
    if B.size < A.rows * B_rows:
        raise ValueError("B must hold B_rows * rows floats (column-major)")
    Cbuf = np.zeros((A.cols, B_rows), dtype=np.float32)
    dt = C.c_float(0)
    rp, nz, jab = (np.ascontiguousarray(x, dtype=np.int64) for x in (A.row_part, A.nzcount, A.jab))
    mab = np.ascontiguousarray(A.mab, dtype=np.float32)
    prec = _lib.PRECISIONS[precision]
    _lib._check(lib.sparta_vbr_spmm_BA(A.rows, A.cols, A.block_rows, A.block_col_size, _lib._ptr(rp),
                                       _lib._ptr(nz), _lib._ptr(jab), _lib._ptr(mab), _lib._ptr(B),
                                       B_rows, B_rows, _lib._ptr(Cbuf), B_rows, prec, C.byref(dt)))
    return Cbuf, dt.value


def bellpack_from_vbr(A: VBR):
    """Host repack VBR -> Blocked-ELL with the exact output of
    prepare_cusparse_BLOCKEDELLPACK (reference cuda_utilities.cpp:1656-1710).
    Returns (ell_blocksize, ellColInd[int64 rows/bs x max_nz], ellValues[rows x max_nz*bs])."""
    bs = A.block_col_size
    if A.rows // bs != A.block_rows or np.any(np.diff(A.row_part) != bs):
        raise ValueError("Blocked-ELL needs square fixed-size blocks")
    return _lib.host_bellpack_from_vbr(A.rows, A.cols, bs, A.nzcount, A.jab, A.mab)


def bellpack_spmm(rows, cols, blocksize, ell_col_ind, ell_values, B, B_cols, precision="bf16"):
    """C = A*B like cusparse_gemm_custom_ellpack / compute_cutlass_bellpack (reference
    cuda_utilities.cpp:1497, cutlass_bellpack_lib.cu:61): B and C row-major.  Returns (C, dt_ms)."""
    lib = _lib.load()
    ind = np.ascontiguousarray(ell_col_ind, dtype=np.int64)
    vals = np.ascontiguousarray(ell_values, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    Cbuf = np.zeros((rows, B_cols), dtype=np.float32)
    dt = C.c_float(0)
    prec = _lib.PRECISIONS[precision]
    _lib._check(lib.sparta_bellpack_spmm(rows, cols, blocksize, ind.shape[0], ind.shape[1],
                                         _lib._ptr(ind), _lib._ptr(vals), _lib._ptr(B), B_cols,
                                         B_cols, _lib._ptr(Cbuf), B_cols, prec, C.byref(dt)))
    return Cbuf, dt.value


def csr_spmm(rows, cols, rowptr, colind, val, B, B_cols, precision="bf16"):
    """C = A*B like cusparse_blockmat_multiplyAB (reference cuda_utilities.cpp:1479-1493, `-M 2`):
    A as the flat CSR prepare_cusparse_CSR builds (val=None for a pattern-only matrix), B and C
    row-major.  precision "tf32" runs plain fp32 (bit-identical to CSR::multiply).  Returns (C, dt_ms)."""
    lib = _lib.load()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colind = np.ascontiguousarray(colind, dtype=np.int64)
    val = None if val is None else np.ascontiguousarray(val, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    Cbuf = np.zeros((rows, B_cols), dtype=np.float32)
    dt = C.c_float(0)
    prec = _lib.PRECISIONS[precision]
    _lib._check(lib.sparta_csr_spmm(rows, cols, _lib._ptr(rowptr), _lib._ptr(colind),
                                    None if val is None else _lib._ptr(val), _lib._ptr(B), B_cols, B_cols,
                                    _lib._ptr(Cbuf), B_cols, prec, C.byref(dt)))
    return Cbuf, dt.value
