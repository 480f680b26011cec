"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU checkers.

  * `Oracle()`     -> oracle/libsparta_oracle.so, the restatement (sparta_oracle.cpp)
  * `Reference()`  -> oracle/_ref/libsparta_ref.so, the unmodified reference sources
                      behind oracle/ref_driver.cpp (None when it was never built)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  Nothing in sparta_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libsparta_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsparta_ref.so")


class Result(C.Structure):
    """Layout shared by OracleResult (sparta_oracle.cpp) and RefResult (ref_driver.cpp)."""
    _fields_ = [
        ("csr_rows", C.c_long), ("csr_cols", C.c_long), ("csr_nnz", C.c_long),
        ("csr_rowptr", C.POINTER(C.c_long)), ("csr_colind", C.POINTER(C.c_long)),
        ("csr_val", C.POINTER(C.c_float)), ("grouping", C.POINTER(C.c_long)),
        ("comparison_counter", C.c_long), ("merge_counter", C.c_long),
        ("average_merge_tau", C.c_float), ("average_row_distance", C.c_float),
        ("VBR_nzcount", C.c_long), ("VBR_nzblocks_count", C.c_long), ("VBR_longest_row", C.c_long),
        ("VBR_average_height", C.c_float),
        ("rows", C.c_long), ("cols", C.c_long), ("block_rows", C.c_long), ("block_cols", C.c_long),
        ("block_col_size", C.c_long), ("nztot", C.c_long), ("jab_len", C.c_long),
        ("row_part", C.POINTER(C.c_long)), ("nzcount", C.POINTER(C.c_long)),
        ("jab", C.POINTER(C.c_long)), ("mab", C.POINTER(C.c_float)),
    ]


def _arr(ptr, n, dtype):
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _unpack(r, fill):
    out = {
        "csr_rows": r.csr_rows, "csr_cols": r.csr_cols, "csr_nnz": r.csr_nnz,
        "csr_rowptr": _arr(r.csr_rowptr, r.csr_rows + 1, np.int64),
        "csr_colind": _arr(r.csr_colind, r.csr_nnz, np.int64),
        "csr_val": _arr(r.csr_val, r.csr_nnz, np.float32),
        "grouping": _arr(r.grouping, r.csr_rows, np.int64),
        "comparison_counter": r.comparison_counter, "merge_counter": r.merge_counter,
        "average_merge_tau": r.average_merge_tau, "average_row_distance": r.average_row_distance,
        "VBR_nzcount": r.VBR_nzcount, "VBR_nzblocks_count": r.VBR_nzblocks_count,
        "VBR_longest_row": r.VBR_longest_row, "VBR_average_height": r.VBR_average_height,
    }
    if fill:
        out.update({
            "rows": r.rows, "cols": r.cols, "block_rows": r.block_rows, "block_cols": r.block_cols,
            "block_col_size": r.block_col_size, "nztot": r.nztot,
            "row_part": _arr(r.row_part, r.block_rows + 1, np.int64),
            "nzcount": _arr(r.nzcount, r.block_rows, np.int64),
            "jab": _arr(r.jab, r.jab_len, np.int64),
            "mab": _arr(r.mab, r.nztot, np.float32),
        })
    return out


def build(verbose=False):
    """Compile the restatement (always) and oracle/_ref (only where /root/reference exists)."""
    res = subprocess.run(["make", "-C", HERE], capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + (res.stdout or "") + (res.stderr or ""))


def _l(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _slack(Bm, w):
    """VBR::multiply walks k < w in the last column block even when cols % w != 0 (vbr.cpp:362),
    reading up to w-1 floats past the end of B's last column (the matching A entries are zero).
    Append w finite zeros so that 0 * whatever-follows-the-buffer cannot turn into NaN."""
    flat = _f(Bm).reshape(-1)
    return np.concatenate([flat, np.zeros(w, dtype=np.float32)])


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.oracle_run.restype = C.c_int
        L.oracle_run.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_float, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.POINTER(Result)]
        L.oracle_free.argtypes = [C.POINTER(Result)]
        L.oracle_vbr_multiply.restype = None
        L.oracle_vbr_multiply.argtypes = [C.c_long, C.c_long] + [C.c_void_p] * 5 + [C.c_long, C.c_long,
                                                                                     C.c_void_p, C.c_long]
        L.oracle_vbr_multiply_BA.restype = None
        L.oracle_vbr_multiply_BA.argtypes = [C.c_long, C.c_long, C.c_long] + [C.c_void_p] * 5 + [
            C.c_long, C.c_long, C.c_void_p, C.c_long]
        L.oracle_csr_multiply.restype = None
        L.oracle_csr_multiply.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_distance.restype = C.c_float
        L.oracle_distance.argtypes = [C.c_int, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_long,
                                      C.c_long, C.c_long]
        L.oracle_merge_rows.restype = C.c_long
        L.oracle_merge_rows.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_permutation.argtypes = [C.c_long, C.c_void_p, C.c_void_p]
        L.oracle_partition.restype = C.c_long
        L.oracle_partition.argtypes = [C.c_long, C.c_void_p, C.c_void_p]
        L.oracle_vbr_fill.restype = C.c_int
        L.oracle_vbr_fill.argtypes = [C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_long, C.c_long, C.c_int, C.POINTER(Result)]
        L.oracle_bellpack_width.restype = C.c_long
        L.oracle_bellpack_width.argtypes = [C.c_long, C.c_void_p]
        L.oracle_bellpack_from_vbr.restype = C.c_long
        L.oracle_bellpack_from_vbr.argtypes = [C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p]

    def run(self, path, a=3, b=3, B=3, t=0.1, m=1, P=0, r=0, s=0, F=0, g=0, p=1, e=0, delim=" ",
            fill=True):
        """Flags named after the reference CLI (include/input.h:81)."""
        res = Result()
        rc = self.lib.oracle_run(path.encode(), delim.encode(), P, e, r, s, a, m, t, b, B, g, p, F,
                                 int(fill), C.byref(res))
        if rc:
            raise RuntimeError(f"oracle_run failed ({rc})")
        try:
            return _unpack(res, fill)
        finally:
            self.lib.oracle_free(C.byref(res))

    def vbr_multiply(self, v, Bm, n, ldb=None, C_init=None):
        """VBR::multiply restatement; returns C as [n, rows] (row j = column j of C)."""
        rows, cols = int(v["rows"]), int(v["cols"])
        ldb = cols if ldb is None else ldb
        Bm = _slack(Bm, int(v["block_col_size"]))
        Cm = np.zeros(n * rows, dtype=np.float32) if C_init is None else _f(C_init).reshape(-1).copy()
        rp, nz, jab, mab = _l(v["row_part"]), _l(v["nzcount"]), _l(v["jab"]), _f(v["mab"])
        self.lib.oracle_vbr_multiply(len(nz), int(v["block_col_size"]), _p(rp), _p(nz), _p(jab), _p(mab),
                                     _p(Bm), ldb, n, _p(Cm), rows)
        return Cm.reshape(n, rows)

    def vbr_multiply_BA(self, v, Bt, m):
        """C = B*A (the arithmetic cublas_blockmat_multiplyBA intends; parity unpinned, see the C
        source).  Bt: [rows, m] (row k = column k of B); returns C as [cols, m]."""
        rows, cols = int(v["rows"]), int(v["cols"])
        Bt = _f(Bt).reshape(-1)
        Cm = np.zeros(m * cols, dtype=np.float32)
        rp, nz, jab, mab = _l(v["row_part"]), _l(v["nzcount"]), _l(v["jab"]), _f(v["mab"])
        self.lib.oracle_vbr_multiply_BA(cols, len(nz), int(v["block_col_size"]), _p(rp), _p(nz), _p(jab),
                                        _p(mab), _p(Bt), m, m, _p(Cm), m)
        return Cm.reshape(cols, m)

    def ref_multiplyBA_literal(self, v, Bt, m):
        """cublas_blockmat_multiplyBA (cuda_utilities.cpp:640-690) index for index, GEMMs done in
        fp32 on the CPU.  Bt: [rows, m] (row k = column k of the m x rows column-major B).
        Returns (C as [cols_padded, m], out_of_range flag)."""
        w = int(v["block_col_size"])
        cols_pad = ((int(v["cols"]) - 1) // w + 1) * w
        Bt = _f(Bt).reshape(-1)
        Cm = np.zeros(m * cols_pad, dtype=np.float32)
        rp, nz, jab, mab = _l(v["row_part"]), _l(v["nzcount"]), _l(v["jab"]), _f(v["mab"])
        self.lib.oracle_ref_multiplyBA_literal.restype = C.c_int
        self.lib.oracle_ref_multiplyBA_literal.argtypes = [C.c_long, C.c_long] + [C.c_void_p] * 5 + \
            [C.c_long, C.c_long, C.c_void_p]
        oor = self.lib.oracle_ref_multiplyBA_literal(len(nz), w, _p(rp), _p(nz), _p(jab), _p(mab), _p(Bt),
                                                     len(Bt), m, _p(Cm))
        return Cm.reshape(cols_pad, m), bool(oor)

    def csr_multiply(self, rows, rowptr, colind, val, pattern_only, Bm, n):
        rowptr, colind, val, Bm = _l(rowptr), _l(colind), _f(val), _f(Bm).reshape(-1)
        Cm = np.zeros(n * rows, dtype=np.float32)
        self.lib.oracle_csr_multiply(rows, _p(rowptr), _p(colind), _p(val), int(pattern_only), _p(Bm), n,
                                     _p(Cm))
        return Cm.reshape(n, rows)

    def distance(self, measure, a, ga, b, gb, w):
        a, b = _l(a), _l(b)
        return float(self.lib.oracle_distance(measure, _p(a), len(a), ga, _p(b), len(b), gb, w))

    def merge_rows(self, a, b):
        a, b = _l(a), _l(b)
        out = np.zeros(len(a) + len(b) + 1, dtype=np.int64)
        n = self.lib.oracle_merge_rows(_p(a), len(a), _p(b), len(b), _p(out))
        return out[:n]

    def permutation(self, grouping):
        g = _l(grouping)
        out = np.zeros(len(g), dtype=np.int64)
        self.lib.oracle_permutation(len(g), _p(g), _p(out))
        return out

    def partition(self, grouping):
        g = _l(grouping)
        out = np.zeros(len(g) + 1, dtype=np.int64)
        n = self.lib.oracle_partition(len(g), _p(g), _p(out))
        return out[:n]

    def vbr_fill(self, rows, cols, rowptr, colind, val, pattern_only, grouping, w, row_block_size=0,
                 force_fixed=False):
        res = Result()
        rowptr, colind, grouping = _l(rowptr), _l(colind), _l(grouping)
        val = _f(val) if val is not None else np.ones(len(colind), dtype=np.float32)
        self.lib.oracle_vbr_fill(rows, cols, _p(rowptr), _p(colind), _p(val), int(pattern_only),
                                 _p(grouping), w, row_block_size, int(force_fixed), C.byref(res))
        try:
            res.csr_rows = 0
            out = _unpack(res, True)
        finally:
            self.lib.oracle_free(C.byref(res))
        return {k: out[k] for k in ("rows", "cols", "block_rows", "block_cols", "block_col_size",
                                    "nztot", "row_part", "nzcount", "jab", "mab")}

    def bellpack_from_vbr(self, v):
        rows, cols, bs = int(v["rows"]), int(v["cols"]), int(v["block_col_size"])
        nz, jab, mab = _l(v["nzcount"]), _l(v["jab"]), _f(v["mab"])
        width = self.lib.oracle_bellpack_width(rows // bs, _p(nz))
        ind = np.zeros((rows // bs, width), dtype=np.int64)
        vals = np.zeros((rows, width * bs), dtype=np.float32)
        w2 = self.lib.oracle_bellpack_from_vbr(rows, cols, bs, _p(nz), _p(jab), _p(mab), _p(ind), _p(vals))
        if w2 < 0:
            raise ValueError("rows/cols not multiples of the block size")
        return bs, ind, vals


class Reference:
    """The unmodified reference behind ref_driver.cpp; `Reference.available()` gates its use."""

    @staticmethod
    def available(variant=None):
        return os.path.exists(Reference.path(variant))

    @staticmethod
    def path(variant=None):
        """variant None: the -O2 build; "O0" / "O3": the reference's own optimisation levels
        (`make serial` has no -O flag, makefile:2; the cluster makefiles use -O3)."""
        return REF_SO if not variant else REF_SO.replace(".so", f"_{variant}.so")

    def __init__(self, variant=None):
        self.lib = C.CDLL(Reference.path(variant))
        L = self.lib
        L.ref_run.restype = C.c_int
        L.ref_run.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.POINTER(Result)]
        L.ref_free.argtypes = [C.POINTER(Result)]
        L.ref_vbr_multiply.restype = None
        L.ref_vbr_multiply.argtypes = [C.c_long] * 4 + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
        L.ref_hamming.restype = C.c_float
        L.ref_jaccard.restype = C.c_float
        for fn in (L.ref_hamming, L.ref_jaccard):
            fn.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_long, C.c_long, C.c_long]

    def run(self, path, fill=True, **flags):
        """flags: reference CLI letters, e.g. a=5, b=64, B=64, t=0.6, P=1."""
        argv = ["ref", "-f", path, "-v", "0"]
        for k, v in flags.items():
            argv += [f"-{k}", str(v)]
        arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
        res = Result()
        rc = self.lib.ref_run(len(argv), arr, int(fill), C.byref(res))
        if rc:
            raise RuntimeError("ref_run failed")
        try:
            return _unpack(res, fill)
        finally:
            self.lib.ref_free(C.byref(res))

    def vbr_multiply(self, v, Bm, n):
        rows, cols = int(v["rows"]), int(v["cols"])
        Bm = _slack(Bm, int(v["block_col_size"]))
        Cm = np.zeros(n * rows, dtype=np.float32)
        rp, nz, jab, mab = _l(v["row_part"]), _l(v["nzcount"]), _l(v["jab"]), _f(v["mab"])
        self.lib.ref_vbr_multiply(rows, cols, len(nz), int(v["block_col_size"]), _p(rp), _p(nz), _p(jab),
                                  _p(mab), _p(Bm), n, _p(Cm))
        return Cm.reshape(n, rows)

    def distance(self, measure, a, ga, b, gb, w):
        a, b = _l(a), _l(b)
        fn = self.lib.ref_hamming if measure == 0 else self.lib.ref_jaccard
        return float(fn(_p(a), len(a), ga, _p(b), len(b), gb, w))
