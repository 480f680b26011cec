"""world_size-2 gloo run of the multi-GPU host logic (sparta_b200/dist.py): shard ranges,
the B broadcast and the ragged all-gather of C.  The per-shard product is computed by the oracle
here (CPU box); on the GPU box the same flow runs in bench.py with the CUDA kernel."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.oracle_py import Oracle
    from sparta_b200 import dist as sd
    from tests.util import random_vbr

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)                       # same matrix on every rank
    heights = rng.integers(1, 70, size=23)
    v = random_vbr(rng, 23, 300, 32, heights, 0.4, values="int")
    n = 17
    Bm = rng.integers(-3, 4, size=(n, 300)).astype(np.float32) if rank == 0 else None
    if os.environ.get("SPARTA_TEST_PARTITION") == "model":
        lo, hi, cuts = sd.shard_range(v["row_part"], v["nzcount"], world, rank, jab=v["jab"], cols=300,
                                      block_col_size=32, n=n)
    else:
        lo, hi, cuts = sd.shard_range(v["row_part"], v["nzcount"], world, rank)
    assert cuts[0] == 0 and cuts[-1] == 23 and np.all(np.diff(cuts) >= 0)
    Bd = sd.broadcast_B(Bm, (n, 300), torch.device("cpu"))
    # the shard's product (stand-in for the CUDA kernel on this CPU box)
    jab_off = np.concatenate([[0], np.cumsum(v["nzcount"])])
    mab_off = np.concatenate([[0], np.cumsum(v["nzcount"] * np.diff(v["row_part"]) * 32)])
    sub = {"rows": sd.shard_rows(v["row_part"], lo, hi), "cols": 300, "block_col_size": 32,
           "row_part": v["row_part"][lo:hi + 1] - v["row_part"][lo], "nzcount": v["nzcount"][lo:hi],
           "jab": v["jab"][jab_off[lo]:jab_off[hi]], "mab": v["mab"][mab_off[lo]:mab_off[hi]]}
    slab = torch.from_numpy(Oracle().vbr_multiply(sub, Bd.numpy(), n)) if sub["rows"] else torch.zeros((n, 0))
    rows_per_rank = [sd.shard_rows(v["row_part"], int(cuts[r]), int(cuts[r + 1])) for r in range(world)]
    full = sd.all_gather_C(slab, rows_per_rank, n)
    ref = Oracle().vbr_multiply(v, Bd.numpy(), n)
    ok = np.array_equal(full.numpy(), ref)
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("partition", ["area", "model"])
def test_two_rank_shard_broadcast_gather(tmp_path, oracle, partition, monkeypatch):
    monkeypatch.setenv("SPARTA_TEST_PARTITION", partition)
    port = 29500 + (os.getpid() + (7 if partition == "model" else 0)) % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"


def test_partition_balances_area(lib):
    import sparta_b200
    rng = np.random.default_rng(1)
    heights = rng.integers(1, 200, size=500)
    row_part = np.concatenate([[0], np.cumsum(heights)])
    nz = rng.integers(0, 60, size=500)
    area = nz * heights
    for parts in (2, 4, 8):
        cuts = sparta_b200.partition_block_rows(row_part, nz, parts)
        loads = [area[cuts[i]:cuts[i + 1]].sum() for i in range(parts)]
        assert max(loads) <= 1.15 * (area.sum() / parts) + area.max()
    cuts = sparta_b200.partition_block_rows(row_part[:4], nz[:3], 8)   # more ranks than block-rows
    assert cuts[0] == 0 and cuts[-1] == 3 and np.all(np.diff(cuts) >= 0)
