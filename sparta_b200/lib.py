"""ctypes binding of the C ABI declared in include/sparta_b200.h.

The shared library is built in-tree (sparta_b200/build.py) and loaded from
sparta_b200/libsparta_b200.so.  There is no Python or CPU fallback: if the
library is missing or a compute entry point fails, the caller gets an exception.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SPARTA_B200_LIB selects another build of the same ABI (A/B runs of two kernel versions on one box)
LIB_PATH = os.environ.get("SPARTA_B200_LIB") or os.path.join(HERE, "libsparta_b200.so")

BF16, FP16, TF32 = 0, 1, 2
LAYOUT_DEFAULT, COL_MAJOR, ROW_MAJOR = 0, 1, 2
PRECISIONS = {"bf16": BF16, "fp16": FP16, "tf32": TF32}


class SpartaError(RuntimeError):
    pass


class Options(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32),
        ("b_layout", C.c_int32), ("c_layout", C.c_int32), ("accumulate", C.c_int32),
        ("seg_rows", C.c_int32), ("acc_cols", C.c_int32), ("panel_stages", C.c_int32),
        ("num_ctas", C.c_int32), ("block_row_begin", C.c_int64), ("block_row_end", C.c_int64),
        ("cta_pair", C.c_int32), ("row_order", C.c_int32), ("l2_slab_mb", C.c_int32),
        ("max_chain", C.c_int32), ("split_k", C.c_int32), ("fuse_rows", C.c_int32),
        ("explicit_range", C.c_int32), ("pipeline", C.c_int32), ("copy_warps", C.c_int32), ("gather_max_height", C.c_int32), ("gather_passes", C.c_int32), ("wide_tiles", C.c_int32), ("n_hint", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("cols", C.c_int64), ("block_rows", C.c_int64),
        ("nz_blocks", C.c_int64), ("nztot", C.c_int64), ("segments", C.c_int64),
        ("super_rows", C.c_int64), ("chunks", C.c_int64), ("items", C.c_int64),
        ("a_packed_bytes", C.c_int64), ("b_bytes", C.c_int64), ("c_bytes", C.c_int64),
        ("grid", C.c_int32), ("smem_bytes", C.c_int32), ("sched_imbalance", C.c_double),
        ("upload_ms", C.c_double), ("kernel_launches", C.c_int64),
        ("team", C.c_int32), ("cta_pair", C.c_int32), ("split_pieces", C.c_int32), ("zero_tiles", C.c_int32),
        ("sched_max_cycles", C.c_double), ("gather_rows", C.c_int64), ("gather_nnz", C.c_int64),
        ("wide_tiles", C.c_int32), ("reserved3", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/sparta_b200.h declares
SIGNATURES = {
    "sparta_last_error": (C.c_char_p, []),
    "sparta_abi_version": (C.c_int, []),
    "sparta_device_count": (C.c_int, []),
    "sparta_vbr_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                    _vp, _vp, _vp, _vp, C.POINTER(Options)]),
    "sparta_vbr_create_from_csr": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64,
                                             C.c_int64, C.c_int32, C.POINTER(Options), _vp]),
    "sparta_csr_vbr_spmm": (C.c_int, [C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int32,
                                      _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int, C.POINTER(C.c_float)]),
    "sparta_vbr_create_BA": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                       _vp, _vp, _vp, _vp, C.POINTER(Options)]),
    "sparta_bellpack_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                         C.c_int64, _vp, _vp, C.POINTER(Options)]),
    "sparta_csr_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, _vp, _vp, _vp,
                                    C.POINTER(Options)]),
    "sparta_set_B": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, C.c_int]),
    "sparta_set_C": (C.c_int, [_vp, _vp, C.c_int64, C.c_int]),
    "sparta_run": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "sparta_run_async": (C.c_int, [_vp]),
    "sparta_run_traced": (C.c_int, [_vp, C.c_int32, _vp, C.c_int64]),
    "sparta_synchronize": (C.c_int, [_vp]),
    "sparta_get_C": (C.c_int, [_vp, _vp, C.c_int64, C.c_int]),
    "sparta_get_C_permuted": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, C.c_int]),
    "sparta_C_device_ptr": (_vp, [_vp]),
    "sparta_C_device_ld": (C.c_int64, [_vp]),
    "sparta_stream": (_vp, [_vp]),
    "sparta_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "sparta_destroy": (C.c_int, [_vp]),
    "sparta_vbr_spmm": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp,
                                  _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int,
                                  C.POINTER(C.c_float)]),
    "sparta_vbr_spmm_multi": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp,
                                        _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int, C.c_int32, C.c_int32,
                                        C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "sparta_vbr_spmm_BA": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp,
                                     _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int,
                                     C.POINTER(C.c_float)]),
    "sparta_bellpack_spmm": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp,
                                       _vp, _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int,
                                       C.POINTER(C.c_float)]),
    "sparta_csr_spmm": (C.c_int, [C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, _vp,
                                  C.c_int64, C.c_int, C.POINTER(C.c_float)]),
    "sparta_release_workspace": (C.c_int, []),
    "sparta_partition_block_rows": (C.c_int, [C.c_int64, _vp, _vp, C.c_int32, _vp]),
    "sparta_partition_block_rows_modelled": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp,
                                                       C.c_int64, C.POINTER(Options), C.c_int32, _vp]),
    "sparta_partition_model_times": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp,
                                               C.c_int64, C.POINTER(Options), C.c_int32, _vp, _vp]),
    "sparta_partition_block_rows_measured": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp,
                                                       C.c_int64, C.POINTER(Options), C.c_int32, _vp, _vp]),
    "sparta_host_blocking": (C.c_int, [C.c_int64, C.c_int64, _vp, _vp, C.c_int32, C.c_float, C.c_int64,
                                       C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       _vp, _vp]),
    "sparta_grouping_save": (C.c_int, [C.c_char_p, C.c_int64, _vp, C.c_uint64]),
    "sparta_grouping_load": (C.c_int, [C.c_char_p, C.c_int64, _vp, C.c_uint64]),
    "sparta_blocking_key": (C.c_uint64, [C.c_int64, C.c_int64, _vp, _vp, C.c_int32, C.c_float, C.c_int64,
                                         C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "sparta_host_blocking_cached": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, _vp, _vp, C.c_int32, C.c_float,
                                              C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_int32, _vp, _vp, C.POINTER(C.c_int32)]),
    "sparta_host_row_order": (C.c_int, [C.c_int64, _vp, C.c_int32, C.c_uint32, _vp]),
    "sparta_host_permutation": (C.c_int, [C.c_int64, _vp, _vp]),
    "sparta_host_partition": (C.c_int, [C.c_int64, _vp, _vp, C.POINTER(C.c_int64)]),
    "sparta_host_vbr_fill": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, _vp, _vp, _vp, C.c_int32,
                                       _vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32]),
    "sparta_host_vbr_get": (C.c_int, [_vp, _vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                      C.POINTER(_vp)]),
    "sparta_host_vbr_free": (C.c_int, [_vp]),
    "sparta_host_bellpack_from_vbr": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, _vp,
                                                _vp, _vp, C.c_int32]),
    "sparta_host_bellpack_get": (C.c_int, [_vp, _vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "sparta_host_bellpack_free": (C.c_int, [_vp]),
    "sparta_vbr_plan_create": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                         _vp, _vp, _vp, C.c_int64, C.POINTER(Options)]),
    "sparta_vbr_plan_create_BA": (C.c_int, [C.POINTER(_vp), C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                            _vp, _vp, _vp, C.c_int64, C.POINTER(Options)]),
    "sparta_plan_array": (C.c_int, [_vp, C.c_int32, C.POINTER(_vp), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32)]),
    "sparta_plan_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "sparta_plan_destroy": (C.c_int, [_vp]),
}

_lib = None


def load():
    """Load libsparta_b200.so (raises SpartaError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SpartaError(
            f"{LIB_PATH} is missing: build it with `python sparta_b200/build.py` "
            "(there is no fallback implementation)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise SpartaError(f"[{rc}] {load().sparta_last_error().decode()}")


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None or a.size == 0 else a.ctypes.data_as(_vp)


def make_options(precision="bf16", **kw):
    o = Options()
    o.struct_size = C.sizeof(Options)
    o.precision = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    for k, v in kw.items():
        if v is None:
            continue
        if k == "device":
            v = int(v) + 1  # ABI: 1 + ordinal, 0 = current
        setattr(o, k, int(v))
    # a range passed from Python is literal: [0, 0) is an empty shard, not "everything"
    if kw.get("block_row_end") is not None:
        o.explicit_range = 1
    return o


def partition_block_rows(row_part, nzcount, parts):
    row_part, nzcount = _i64(row_part), _i64(nzcount)
    cuts = np.zeros(parts + 1, dtype=np.int64)
    _check(load().sparta_partition_block_rows(len(nzcount), _ptr(row_part), _ptr(nzcount), parts,
                                              _ptr(cuts)))
    return cuts


def partition_block_rows_modelled(rows, cols, block_col_size, row_part, nzcount, jab, n, parts, **opts):
    """Contiguous block-row ranges balanced on the scheduler's modelled shard times."""
    row_part, nzcount, jab = _i64(row_part), _i64(nzcount), _i64(jab)
    cuts = np.zeros(parts + 1, dtype=np.int64)
    o = make_options(**opts)
    _check(load().sparta_partition_block_rows_modelled(rows, cols, len(nzcount), block_col_size, _ptr(row_part),
                                                       _ptr(nzcount), _ptr(jab), n, C.byref(o), parts, _ptr(cuts)))
    return cuts


def partition_model_times(rows, cols, block_col_size, row_part, nzcount, jab, n, cuts, **opts):
    """The scheduler model's time (SM cycles) of every shard of the partition `cuts`."""
    row_part, nzcount, jab, cuts = _i64(row_part), _i64(nzcount), _i64(jab), _i64(cuts)
    out = np.zeros(len(cuts) - 1, dtype=np.float64)
    o = make_options(**opts)
    _check(load().sparta_partition_model_times(rows, cols, len(nzcount), block_col_size, _ptr(row_part), _ptr(nzcount),
                                               _ptr(jab), n, C.byref(o), len(cuts) - 1, _ptr(cuts),
                                               out.ctypes.data_as(C.c_void_p)))
    return out


def partition_block_rows_measured(rows, cols, block_col_size, row_part, nzcount, jab, n, parts, prev_cuts,
                                  measured_ms, modelled_cycles, **opts):
    """The modelled partition corrected by measured shard times of an earlier partition `prev_cuts`:
    every block-row gets the measured / modelled ratio of the shard it was in (ratios are
    normalised to mean 1, so the clock rate does not matter)."""
    row_part, nzcount, jab = _i64(row_part), _i64(nzcount), _i64(jab)
    ratio = np.asarray(measured_ms, dtype=np.float64) / np.maximum(np.asarray(modelled_cycles, dtype=np.float64), 1.0)
    ratio = ratio / ratio.mean()
    scale = np.ones(len(nzcount), dtype=np.float64)
    for r in range(len(prev_cuts) - 1):
        scale[int(prev_cuts[r]):int(prev_cuts[r + 1])] = ratio[r]
    cuts = np.zeros(parts + 1, dtype=np.int64)
    o = make_options(**opts)
    _check(load().sparta_partition_block_rows_measured(rows, cols, len(nzcount), block_col_size, _ptr(row_part),
                                                       _ptr(nzcount), _ptr(jab), n, C.byref(o), parts,
                                                       scale.ctypes.data_as(C.c_void_p), _ptr(cuts)))
    return cuts


class Handle:
    """Device-resident A (VBR or Blocked-ELL) plus the current B/C buffers."""

    def __init__(self, ptr, keep):
        self._h = ptr
        self._keep = keep

    @classmethod
    def from_vbr(cls, rows, cols, block_col_size, row_part, nzcount, jab, mab, **opts):
        lib = load()
        row_part, nzcount, jab, mab = _i64(row_part), _i64(nzcount), _i64(jab), _f32(mab)
        o = make_options(**opts)
        h = _vp()
        _check(lib.sparta_vbr_create(C.byref(h), rows, cols, len(nzcount), block_col_size,
                                     _ptr(row_part), _ptr(nzcount), _ptr(jab), _ptr(mab),
                                     C.byref(o)))
        return cls(h, None)

    @classmethod
    def from_csr_grouping(cls, rows, cols, rowptr, colind, val, grouping, block_col_size, row_block_size=0,
                          force_fixed_size=False, **opts):
        """A from the flat CSR and the row grouping (sparta_vbr_create_from_csr): the dense blocks are
        rebuilt on the device.  Returns the handle; handle.vbr_dims = rows, cols, block_rows, block_cols,
        block_col_size, nztot of the VBR."""
        lib = load()
        rowptr, colind, grouping = _i64(rowptr), _i64(colind), _i64(grouping)
        val = None if val is None else _f32(val)
        o = make_options(**opts)
        h = _vp()
        dims = np.zeros(6, dtype=np.int64)
        _check(lib.sparta_vbr_create_from_csr(C.byref(h), rows, cols, _ptr(rowptr), _ptr(colind),
                                              None if val is None else _ptr(val), _ptr(grouping), block_col_size,
                                              row_block_size, int(force_fixed_size), C.byref(o), _ptr(dims)))
        obj = cls(h, None)
        obj.vbr_dims = dims
        return obj

    @classmethod
    def from_vbr_BA(cls, rows, cols, block_col_size, row_part, nzcount, jab, mab, **opts):
        """The inverted product C = B*A (-M 6): B is [rows][n], C is [cols][n] (row-major views of
        the reference's column-major n x rows and n x cols arrays)."""
        lib = load()
        row_part, nzcount, jab, mab = _i64(row_part), _i64(nzcount), _i64(jab), _f32(mab)
        o = make_options(**opts)
        h = _vp()
        _check(lib.sparta_vbr_create_BA(C.byref(h), rows, cols, len(nzcount), block_col_size,
                                        _ptr(row_part), _ptr(nzcount), _ptr(jab), _ptr(mab),
                                        C.byref(o)))
        return cls(h, None)

    @classmethod
    def from_bellpack(cls, rows, cols, blocksize, ell_col_ind, ell_values, **opts):
        lib = load()
        ind = _i64(ell_col_ind)
        vals = _f32(ell_values)
        ind_rows, ind_cols = ind.shape
        o = make_options(**opts)
        h = _vp()
        _check(lib.sparta_bellpack_create(C.byref(h), rows, cols, blocksize, ind_rows, ind_cols,
                                          _ptr(ind), _ptr(vals), C.byref(o)))
        return cls(h, None)

    @classmethod
    def from_csr(cls, rows, cols, rowptr, colind, val=None, **opts):
        """Flat CSR (val=None: pattern-only, all ones).  block_row_begin/end select ROWS here."""
        lib = load()
        rowptr, colind = _i64(rowptr), _i64(colind)
        val = None if val is None else _f32(val)
        o = make_options(**opts)
        h = _vp()
        _check(lib.sparta_csr_create(C.byref(h), rows, cols, _ptr(rowptr), _ptr(colind),
                                     None if val is None else _ptr(val), C.byref(o)))
        return cls(h, None)

    def set_B(self, B, ld, n):
        """B: numpy fp32 array (host) in the handle's B layout."""
        B = _f32(B)
        _check(load().sparta_set_B(self._h, _ptr(B), ld, n, 0))

    def set_B_device(self, dptr, ld, n):
        _check(load().sparta_set_B(self._h, _vp(dptr), ld, n, 1))

    def set_C(self, Cbuf, ld):
        Cbuf = _f32(Cbuf)
        _check(load().sparta_set_C(self._h, _ptr(Cbuf), ld, 0))

    def run(self):
        dt = C.c_float(0)
        _check(load().sparta_run(self._h, C.byref(dt)))
        return dt.value

    def run_traced(self, worker=0, capacity=16384):
        """One multiply with worker's timeline: uint64[zone 4][rank 2][capacity][2] SM clocks."""
        rec = np.zeros((4, 2, capacity, 2), dtype=np.uint64)
        _check(load().sparta_run_traced(self._h, worker, _ptr(rec), capacity))
        return rec

    def run_async(self):
        _check(load().sparta_run_async(self._h))

    def synchronize(self):
        _check(load().sparta_synchronize(self._h))

    def get_C(self, out, ld):
        assert out.dtype == np.float32 and out.flags.c_contiguous
        _check(load().sparta_get_C(self._h, _ptr(out), ld, 0))
        return out

    def get_C_permuted(self, out, ld, row_map, out_rows):
        """C with row r of the handle written to row row_map[r] of `out` (original row order);
        `out` has out_rows rows in the handle's C layout."""
        row_map = _i64(row_map)
        assert out.dtype == np.float32 and out.flags.c_contiguous
        _check(load().sparta_get_C_permuted(self._h, _ptr(out), ld, _ptr(row_map), out_rows, 0))
        return out

    def get_C_device(self, dptr, ld):
        _check(load().sparta_get_C(self._h, _vp(dptr), ld, 1))

    @property
    def c_device_ptr(self):
        return load().sparta_C_device_ptr(self._h)

    @property
    def c_device_ld(self):
        return load().sparta_C_device_ld(self._h)

    @property
    def stream(self):
        return load().sparta_stream(self._h)

    def stats(self):
        s = Stats()
        _check(load().sparta_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def close(self):
        if self._h:
            load().sparta_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


SEG_DT = np.dtype([("c_row0", "<i4"), ("h", "<i4"), ("h_pad", "<i4"), ("tmem_col", "<i4")])
SROW_DT = np.dtype([("seg_begin", "<i4"), ("seg_count", "<i4"), ("chunk_begin", "<i4"),
                    ("chunk_count", "<i4"), ("n_cols", "<i4"), ("break_mask", "<u4"), ("pad", "<i4", (2,))])
CHUNK_DT = np.dtype([("k0", "<i4"), ("mask", "<u4"), ("a_off16", "<u4"), ("a_bytes", "<u4"),
                     ("ksteps", "<i4"), ("tbl_bytes", "<u4"), ("tbl_off16", "<u4"),
                     ("pad", "<i4")])
ITEM_DT = np.dtype([("srow", "<i4"), ("j0", "<i4"), ("chunk_off", "<i4"), ("count", "<u4")])
ITEM_NOT_FIRST, ITEM_NOT_LAST, ITEM_ATOMIC, ITEM_COUNT_MASK = 1 << 31, 1 << 30, 1 << 29, (1 << 29) - 1
JOB_DT = np.dtype([("src_base", "<i8"), ("src_rs", "<i8"), ("src_ks", "<i8"), ("h", "<i4"),
                   ("h_pad", "<i4"), ("k_lo", "<i4"), ("k_w", "<i4"), ("dst_off16", "<u4"),
                   ("r_base", "<i4"), ("pad", "<i4", (2,))])
ZERO_DT = np.dtype([("srow", "<i4"), ("j0", "<i4")])
_PLAN_DTYPES = [SEG_DT, SROW_DT, CHUNK_DT, ITEM_DT, np.dtype("<i4"), np.dtype("<i4"), JOB_DT, np.dtype("<u4"), ZERO_DT]
_PLAN_NAMES = ["segs", "srows", "chunks", "items", "cta_ptr", "cta_items", "jobs", "tables", "zero_jobs"]


def vbr_plan(rows, cols, block_col_size, row_part, nzcount, jab, n, transposed=False, **opts):
    """Host-only: the tile schedule the kernel would walk, as numpy record arrays.
    transposed: the schedule of the inverted product C = B*A (Handle.from_vbr_BA)."""
    lib = load()
    row_part, nzcount, jab = _i64(row_part), _i64(nzcount), _i64(jab)
    o = make_options(**opts)
    p = _vp()
    create = lib.sparta_vbr_plan_create_BA if transposed else lib.sparta_vbr_plan_create
    _check(create(C.byref(p), rows, cols, len(nzcount), block_col_size,
                                      _ptr(row_part), _ptr(nzcount), _ptr(jab), n, C.byref(o)))
    try:
        out = {}
        for which, (name, dt) in enumerate(zip(_PLAN_NAMES, _PLAN_DTYPES)):
            data, count, rec = _vp(), C.c_int64(), C.c_int32()
            _check(lib.sparta_plan_array(p, which, C.byref(data), C.byref(count), C.byref(rec)))
            assert rec.value == dt.itemsize, (name, rec.value, dt.itemsize)
            if count.value:
                buf = (C.c_char * (count.value * rec.value)).from_address(data.value)
                out[name] = np.frombuffer(buf, dtype=dt).copy()
            else:
                out[name] = np.zeros(0, dtype=dt)
        s = Stats()
        _check(lib.sparta_plan_stats(p, C.byref(s)))
        out["stats"] = s.as_dict()
        return out
    finally:
        lib.sparta_plan_destroy(p)


def _view(ptr, count, dtype):
    if count <= 0 or not ptr.value:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr.value)
    return np.frombuffer(buf, dtype=dtype).copy()


class BlockingStats(C.Structure):
    _fields_ = [("comparison_counter", C.c_int64), ("merge_counter", C.c_int64),
                ("average_merge_tau", C.c_float), ("average_row_distance", C.c_float),
                ("seconds", C.c_double)]


def host_blocking(rows, cols, rowptr, colind, algo=3, tau=0.1, block_col_size=3, row_block_size=3,
                  sim_measure=1, use_pattern=True, use_group=False, force_fixed_size=False,
                  list_model=False, return_stats=False):
    """BlockingEngine::GetGrouping of the reference (src/general/blocking.cpp:633) on a flat CSR;
    keyword names follow the reference CLI (-a -t -b -B -m -p -g -F)."""
    rowptr, colind = _i64(rowptr), _i64(colind)
    grouping = np.zeros(rows, dtype=np.int64)
    st = BlockingStats()
    _check(load().sparta_host_blocking(rows, cols, _ptr(rowptr), _ptr(colind), int(algo), float(tau),
                                       int(block_col_size), int(row_block_size), int(sim_measure),
                                       int(use_pattern), int(use_group), int(force_fixed_size),
                                       int(list_model), _ptr(grouping), C.addressof(st)))
    if return_stats:
        return grouping, {k: getattr(st, k) for k, _ in st._fields_}
    return grouping


def host_blocking_cached(cache_dir, rows, cols, rowptr, colind, algo=3, tau=0.1, block_col_size=3,
                         row_block_size=3, sim_measure=1, use_pattern=True, use_group=False,
                         force_fixed_size=False):
    """host_blocking behind the library's grouping cache (reference `.g` files + a key sidecar in
    `cache_dir`).  Returns (grouping, hit)."""
    rowptr, colind = _i64(rowptr), _i64(colind)
    grouping = np.zeros(rows, dtype=np.int64)
    hit = C.c_int32(0)
    os.makedirs(cache_dir, exist_ok=True)
    _check(load().sparta_host_blocking_cached(os.fsencode(cache_dir), rows, cols, _ptr(rowptr), _ptr(colind),
                                              int(algo), float(tau), int(block_col_size), int(row_block_size),
                                              int(sim_measure), int(use_pattern), int(use_group),
                                              int(force_fixed_size), 0, _ptr(grouping), None, C.byref(hit)))
    return grouping, bool(hit.value)


def blocking_key(rows, cols, rowptr, colind, algo=3, tau=0.1, block_col_size=3, row_block_size=3, sim_measure=1,
                 use_pattern=True, use_group=False, force_fixed_size=False):
    rowptr, colind = _i64(rowptr), _i64(colind)
    return int(load().sparta_blocking_key(rows, cols, _ptr(rowptr), _ptr(colind), int(algo), float(tau),
                                          int(block_col_size), int(row_block_size), int(sim_measure),
                                          int(use_pattern), int(use_group), int(force_fixed_size)))


def grouping_save(path, grouping, key=0):
    g = _i64(grouping)
    _check(load().sparta_grouping_save(os.fsencode(path), len(g), _ptr(g), int(key)))


def grouping_load(path, rows, key=0):
    g = np.zeros(rows, dtype=np.int64)
    _check(load().sparta_grouping_load(os.fsencode(path), rows, _ptr(g), int(key)))
    return g


def host_row_order(rowptr, mode, seed=0):
    """The reference's -r reordering (1 / -1 degree, 2 scramble with std::rand seeded by -s): order[i] = old
    index of new row i."""
    rowptr = _i64(rowptr)
    order = np.zeros(len(rowptr) - 1, dtype=np.int64)
    _check(load().sparta_host_row_order(len(order), _ptr(rowptr), int(mode), int(seed), _ptr(order)))
    return order


def permute_csr_rows(rowptr, colind, val, order):
    """CSR::permute_rows (src/general/csr.cpp:66-75): new row i = old row order[i]."""
    rowptr, colind = _i64(rowptr), _i64(colind)
    lens = np.diff(rowptr)[order]
    new_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    idx = np.concatenate([np.arange(rowptr[o], rowptr[o + 1]) for o in order]) if len(order) else np.zeros(0, np.int64)
    idx = idx.astype(np.int64)
    return new_ptr, colind[idx], (None if val is None else np.asarray(val)[idx])


def host_permutation(grouping):
    g = _i64(grouping)
    perm = np.zeros(len(g), dtype=np.int64)
    _check(load().sparta_host_permutation(len(g), _ptr(g), _ptr(perm)))
    return perm


def host_partition(grouping):
    g = _i64(grouping)
    part = np.zeros(len(g) + 1, dtype=np.int64)
    n = C.c_int64()
    _check(load().sparta_host_partition(len(g), _ptr(g), _ptr(part), C.byref(n)))
    return part[:n.value]


def host_vbr_fill(rows, cols, rowptr, colind, val, grouping, block_col_size, row_block_size=0,
                  force_fixed_size=False, pattern_only=False, threads=0):
    """Linear-time VBR::fill_from_CSR_inplace; returns a dict with the reference's VBR fields."""
    lib = load()
    rowptr, colind, grouping = _i64(rowptr), _i64(colind), _i64(grouping)
    val = None if val is None else _f32(val)
    if val is None:
        pattern_only = True
    if threads <= 0:
        threads = min(32, os.cpu_count() or 1)
    obj = _vp()
    _check(lib.sparta_host_vbr_fill(C.byref(obj), rows, cols, _ptr(rowptr), _ptr(colind),
                                    None if val is None else _ptr(val), int(pattern_only),
                                    _ptr(grouping), block_col_size, row_block_size,
                                    int(force_fixed_size), threads))
    try:
        dims = np.zeros(6, dtype=np.int64)
        rp, nz, jab, mab = _vp(), _vp(), _vp(), _vp()
        _check(lib.sparta_host_vbr_get(obj, _ptr(dims), C.byref(rp), C.byref(nz), C.byref(jab),
                                       C.byref(mab)))
        br = int(dims[2])
        nzc = _view(nz, br, np.int64)
        return {
            "rows": int(dims[0]), "cols": int(dims[1]), "block_rows": br, "block_cols": int(dims[3]),
            "block_col_size": int(dims[4]), "nztot": int(dims[5]),
            "row_part": _view(rp, br + 1, np.int64), "nzcount": nzc,
            "jab": _view(jab, int(nzc.sum()), np.int64), "mab": _view(mab, int(dims[5]), np.float32),
        }
    finally:
        lib.sparta_host_vbr_free(obj)


def host_bellpack_from_vbr(rows, cols, block_col_size, nzcount, jab, mab, threads=0):
    lib = load()
    nzcount, jab, mab = _i64(nzcount), _i64(jab), _f32(mab)
    if threads <= 0:
        threads = min(32, os.cpu_count() or 1)
    obj = _vp()
    _check(lib.sparta_host_bellpack_from_vbr(C.byref(obj), rows, cols, block_col_size, _ptr(nzcount),
                                             _ptr(jab), _ptr(mab), threads))
    try:
        dims = np.zeros(3, dtype=np.int64)
        ind, vals = _vp(), _vp()
        _check(lib.sparta_host_bellpack_get(obj, _ptr(dims), C.byref(ind), C.byref(vals)))
        bs, ir, ic = (int(x) for x in dims)
        return bs, _view(ind, ir * ic, np.int64).reshape(ir, ic), \
            _view(vals, rows * ic * bs, np.float32).reshape(rows, ic * bs)
    finally:
        lib.sparta_host_bellpack_free(obj)
