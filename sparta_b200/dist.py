"""Multi-GPU plumbing: one process per GPU, `torch.distributed` for the collectives.

The multiply shards by block-row with NO exchange during compute (each block-row reads its own
A blocks plus B and writes its own rows of C, reference src/general/vbr.cpp:342-368), so the
only collectives are one broadcast of B before the multiply and an optional all-gather of the C
row slabs after it.  The backend is whatever the process group was created with:
NCCL over NVLink on the GPU box, gloo in the CPU test-suite.
"""
import numpy as np

from . import lib as _lib


def shard_range(row_part, nzcount, world_size, rank, jab=None, cols=None, block_col_size=None, n=None, **opts):
    """Contiguous block-row range [lo, hi) of `rank`.  With the column-block lists (`jab`, plus
    cols / block_col_size / the number of B columns n) the ranges are balanced on the scheduler's
    modelled kernel time per shard; without them on nonzero-block area."""
    if jab is not None:
        cuts = _lib.partition_block_rows_modelled(int(row_part[-1]), cols, block_col_size, row_part, nzcount, jab,
                                                  n, world_size, **opts)
    else:
        cuts = _lib.partition_block_rows(row_part, nzcount, world_size)
    return int(cuts[rank]), int(cuts[rank + 1]), cuts


def shard_rows(row_part, lo, hi):
    return int(row_part[hi] - row_part[lo])


def broadcast_B(B_host, shape, device, src=0):
    """Replicates the dense operand with ONE broadcast.  `B_host` ([n, cols] fp32, row j =
    column j of B) is only read on rank `src`; every rank gets a device tensor of `shape`."""
    import torch
    import torch.distributed as dist
    t = torch.empty(shape, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(B_host, dtype=np.float32)))
    dist.broadcast(t, src)
    return t


def all_gather_C(C_slab, rows_per_rank, n):
    """Gathers the per-rank C slabs ([n, rows_r] fp32 tensors, blocked row order) into the full
    [n, rows] matrix on every rank.  Slabs are ragged, NCCL has no all-gather-v: every slab is
    padded to the tallest one, gathered, and the padding is cut away."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    tallest = max(int(r) for r in rows_per_rank)
    padded = torch.zeros((n, tallest), dtype=torch.float32, device=C_slab.device)
    if C_slab.numel():
        padded[:, :C_slab.shape[1]] = C_slab
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    return torch.cat([o[:, :int(r)] for o, r in zip(out, rows_per_rank)], dim=1)
