"""C-ABI library: loads, exports every declared symbol, host-only helpers work, and compute entry
points fail loudly (no CPU fallback) when there is no GPU.  CPU only -- no compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sparta_b200
from sparta_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sparta_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sparta_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/sparta_b200.h but not exported"
    assert set(names) == set(L.SIGNATURES), "ctypes binding and header disagree"


def test_abi_version(lib):
    assert lib.sparta_abi_version() == 1


def test_no_oracle_or_cpu_path_in_product():
    """The product must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "sparta_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f == "__init__.py" and False, f"{f} mentions oracle"
    out = os.popen(f"ldd {L.LIB_PATH}").read()
    assert "sparta_oracle" not in out and "sparta_ref" not in out


def test_create_fails_without_gpu_or_runs_on_gpu(lib):
    import torch
    rp = np.array([0, 16], dtype=np.int64)
    nz = np.array([1], dtype=np.int64)
    jab = np.array([0], dtype=np.int64)
    mab = np.ones(16 * 16, dtype=np.float32)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(sparta_b200.SpartaError) as e:
        sparta_b200.Handle.from_vbr(16, 16, 16, rp, nz, jab, mab)
    assert "no CUDA device" in str(e.value) or "[3]" in str(e.value) or "[2]" in str(e.value)
    assert lib.sparta_device_count() == 0


def test_argument_validation(lib):
    h = C.c_void_p()
    rc = lib.sparta_vbr_create(C.byref(h), 16, 16, 1, 0, None, None, None, None, None)
    assert rc == 1 and b"invalid" in lib.sparta_last_error()
    rp = np.array([0, 8], dtype=np.int64)  # row_part[-1] != rows
    nz = np.array([0], dtype=np.int64)
    rc = lib.sparta_vbr_create(C.byref(h), 16, 16, 1, 16, rp.ctypes.data, nz.ctypes.data, None, None, None)
    assert rc == 1 and b"row_part" in lib.sparta_last_error()
    assert lib.sparta_run(None, None) == 1
    assert lib.sparta_set_B(None, None, 0, 0, 0) == 1


def test_partition_block_rows_balances_area():
    rng = np.random.default_rng(0)
    heights = rng.integers(1, 65, size=500)
    row_part = np.concatenate([[0], np.cumsum(heights)])
    nzcount = rng.integers(0, 200, size=500)
    for parts in (1, 2, 4, 8):
        cuts = sparta_b200.partition_block_rows(row_part, nzcount, parts)
        assert cuts[0] == 0 and cuts[-1] == 500 and np.all(np.diff(cuts) >= 0)
        area = nzcount * heights
        loads = [area[cuts[i]:cuts[i + 1]].sum() for i in range(parts)]
        assert max(loads) <= area.sum() / parts + area.max()


def test_partition_skewed_front_loaded():
    # -a 5 puts complete (dense) groups first (blocking.cpp:527-533): equal-count cuts would be unbalanced
    heights = np.full(64, 64)
    row_part = np.concatenate([[0], np.cumsum(heights)])
    nzcount = np.concatenate([np.full(8, 1000), np.full(56, 10)])
    cuts = sparta_b200.partition_block_rows(row_part, nzcount, 4)
    loads = [int((nzcount * heights)[cuts[i]:cuts[i + 1]].sum()) for i in range(4)]
    assert max(loads) / (sum(loads) / 4) < 1.35
