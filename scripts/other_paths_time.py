#!/usr/bin/env python
"""Kernel times of the other entry points on the config #3 matrix (CUDA-event dt of sparta_run):
the inverted product C = B*A (-M 6 / -M 11), the Blocked-ELL route at config #2 (-M 3 / -M 8),
and CSR x dense (-M 2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sparta_b200  # noqa: E402
from sparta_b200 import synth  # noqa: E402


def timed(h, reps=10):
    for _ in range(3):
        h.run()
    return float(np.median([h.run() for _ in range(reps)]))


def main():
    wl = bench.WORKLOADS["rmat16_a5"]
    N, rowptr, colind = bench.make_matrix(wl)
    v = bench.build_vbr(wl, N, rowptr, colind, bench.make_grouping(wl, N, rowptr, colind))
    n = wl["n"]
    flops = 2.0 * v["nztot"] * n
    Bt = synth.seeded_B(v["rows"], n, seed=2).T.copy()          # [rows][n]: column k of B in row k
    for prec in ("bf16", "tf32"):
        h = sparta_b200.Handle.from_vbr_BA(v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"], v["jab"],
                                           v["mab"], precision=prec)
        h.set_B(Bt, n, n)
        ms = timed(h)
        st = h.stats()
        print(f"C = B*A {prec}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s on nonzero-block FLOPs  "
              f"(items {st['items']}, team {st['team']}, split pieces {st['split_pieces']})", flush=True)
        h.close()
    del v
    # config #2 through the Blocked-ELL bundle (row-major B and C)
    wl2 = bench.WORKLOADS["er14_fixed"]
    N2, rp2, ci2 = bench.make_matrix(wl2)
    v2 = bench.build_vbr(wl2, N2, rp2, ci2, bench.make_grouping(wl2, N2, rp2, ci2))
    from sparta_b200.api import VBR, bellpack_from_vbr
    bs, ind, vals = bellpack_from_vbr(VBR(v2["rows"], v2["cols"], wl2["w"], v2["row_part"], v2["nzcount"], v2["jab"], v2["mab"]))
    B2 = synth.seeded_B(v2["cols"], wl2["n"], seed=2).T.copy()  # row-major cols x n
    h = sparta_b200.Handle.from_bellpack(v2["rows"], v2["cols"], bs, ind, vals, precision="bf16")
    h.set_B(B2, wl2["n"], wl2["n"])
    ms = timed(h, 30)
    print(f"Blocked-ELL config #2 bf16: {ms * 1e3:.1f} us  {2.0 * v2['nztot'] * wl2['n'] / ms / 1e9:.1f} TFLOP/s "
          f"(ELL width {ind.shape[1]} blocks, {int((ind >= 0).sum())} real)", flush=True)
    h.close()
    # CSR x dense on the config #3 matrix
    nnz = len(colind)
    B = synth.seeded_B(N, n, seed=2).T.copy()
    for prec in ("bf16", "tf32"):
        h = sparta_b200.Handle.from_csr(N, N, rowptr, colind, None, precision=prec)
        h.set_B(B, n, n)
        ms = timed(h)
        es = 2 if prec == "bf16" else 4
        print(f"CSR {prec if prec == 'bf16' else 'fp32'}: {ms:.3f} ms  nnz={nnz}  gather {nnz * n * es / 1e9:.1f} GB from L2 = "
              f"{nnz * n * es / ms / 1e9:.2f} TB/s; HBM floor {(N * n * es + N * n * 4 + nnz * 8) / 1e9:.2f} GB", flush=True)
        h.close()


if __name__ == "__main__":
    main()
