"""sparta_b200: B200-native (sm_100a) block-sparse x dense multiply behind SPARTA's VBR interface.

Only the hot path lives here: `csrc/` (CUDA kernels + the C ABI of include/sparta_b200.h),
`lib.py` (ctypes binding) and `api.py` (host-side mirror of the reference's multiply entry
points).  Importing the package does not need a GPU; calling a multiply without the built
library or without a CUDA device raises.
"""
from .lib import (BF16, FP16, TF32, COL_MAJOR, ROW_MAJOR, Handle, SpartaError, load,
                  partition_block_rows, partition_block_rows_measured, partition_block_rows_modelled,
                  partition_model_times,
                  vbr_plan)
from .api import VBR, bellpack_from_vbr, bellpack_spmm, csr_spmm, vbr_spmm, vbr_spmm_BA

__all__ = ["BF16", "FP16", "TF32", "COL_MAJOR", "ROW_MAJOR", "Handle", "SpartaError", "load",
           "partition_block_rows", "partition_block_rows_modelled", "partition_block_rows_measured", "partition_model_times", "vbr_plan", "VBR", "vbr_spmm", "bellpack_spmm",
           "bellpack_from_vbr", "csr_spmm", "vbr_spmm_BA"]
