#!/usr/bin/env python
"""BASELINE config #5: a power-law matrix with its rows in the generator's order and scrambled
(what the reference's `-r 2` does: a seeded random row permutation), clustered with
`-a 5 -b 64 -B 64` at tau in {0.0, 0.6, 1.0}, multiplied by B with 256..8192 columns, bf16.
Prints one line per case: nonzero blocks, block density, kernel time, TFLOP/s on nonzero-block
FLOPs, and the sampled fp64 check of bench.py."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sparta_b200  # noqa: E402
from sparta_b200 import lib, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=14)
    ap.add_argument("--density", type=float, default=2e-3)
    ap.add_argument("--ns", default="256,1024,4096,8192")
    ap.add_argument("--taus", default="0.0,0.6,1.0")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    N = 1 << args.scale
    r, c = synth.rmat_edges(args.scale, int(args.density * N * N), seed=1)
    r, c = synth.pin_shape(r, c, N, N)
    results = []
    for order in ("generator", "scrambled"):
        rr = r if order == "generator" else np.random.default_rng(1).permutation(N)[r]
        keep = np.lexsort((c, rr))                      # edges sorted by (row, column)
        rowptr, colind, _ = synth.csr_from_edges(rr[keep], c[keep], N)
        for tau in [float(x) for x in args.taus.split(",")]:
            g = lib.host_blocking(N, N, rowptr, colind, algo=5, tau=tau, block_col_size=64, row_block_size=64,
                                  sim_measure=1, use_pattern=True, use_group=False)
            v = lib.host_vbr_fill(N, N, rowptr, colind, None, g, 64, 64, force_fixed_size=False, pattern_only=True)
            blocks = len(v["jab"])
            dens = blocks / (v["block_rows"] * ((N + 63) // 64))
            for n in [int(x) for x in args.ns.split(",")]:
                Bm = synth.seeded_B(N, n, seed=2)
                h = sparta_b200.Handle.from_vbr(N, N, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"])
                h.set_B(Bm, N, n)
                for _ in range(3):
                    h.run()
                ms = float(np.median([h.run() for _ in range(10)]))
                chk = bench.spot_check(h, v, 0, v["block_rows"], n, "bf16", Bm, None, dev, 0)
                st = h.stats()
                h.close()
                rec = {"rows": order, "tau": tau, "n": n, "nz_blocks": blocks, "block_density": dens,
                       "ms": ms, "tflops": 2.0 * v["nztot"] * n / ms / 1e9, "max_rel_err": chk["max_rel_err"],
                       "ok": chk["ok"], "team": st["team"], "split_pieces": st["split_pieces"]}
                results.append(rec)
                print(f"rows {order:9s} tau {tau:.1f} n {n:5d}: {blocks:6d} blocks ({100 * dens:5.1f} % of the grid)  "
                      f"{ms * 1e3:8.1f} us  {rec['tflops']:7.1f} TFLOP/s  err {chk['max_rel_err']:.1e} ok={chk['ok']}", flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
