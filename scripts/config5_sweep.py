#!/usr/bin/env python
"""BASELINE config #5: a power-law matrix with its rows in the generator's order (-r 0) and scrambled
with the reference's own `-r 2 -s 1` (CSR::scramble = std::random_shuffle driven by std::rand,
src/general/csr.cpp:157-166, through sparta_host_row_order: the same permutation bit for bit),
clustered with `-a 5 -b 64 -B 64` at tau in {0.0, 0.6, 1.0}, swept over BLOCK DENSITY (the element
density of the R-MAT, which sets the share of nonzero blocks) and over B columns 256..8192, bf16.
One line per case: nonzero blocks, block density, kernel time, TFLOP/s on nonzero-block FLOPs and the
sampled fp64 check of bench.py."""
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
    ap.add_argument("--densities", default="5e-4,2e-3,8e-3", help="element densities of the R-MAT (block density follows)")
    ap.add_argument("--ns", default="256,1024,4096,8192")
    ap.add_argument("--taus", default="0.0,0.6,1.0")
    ap.add_argument("--wide-tiles", type=int, default=0, help="sparta_options::wide_tiles (0: the library's choice)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    N = 1 << args.scale
    results = []
    for density in [float(x) for x in args.densities.split(",")]:
        r, c = synth.rmat_edges(args.scale, int(density * N * N), seed=1)
        r, c = synth.pin_shape(r, c, N, N)
        base_ptr, base_col, _ = synth.csr_from_edges(r, c, N)
        for order in ("generator (-r 0)", "scrambled (-r 2 -s 1)"):
            if order.startswith("generator"):
                rowptr, colind = base_ptr, base_col
            else:
                perm = lib.host_row_order(base_ptr, 2, 1)            # new row i = old row perm[i]
                rowptr, colind, _ = lib.permute_csr_rows(base_ptr, base_col, None, perm)
            for tau in [float(x) for x in args.taus.split(",")]:
                g = lib.host_blocking(N, N, rowptr, colind, algo=5, tau=tau, block_col_size=64, row_block_size=64,
                                      sim_measure=1, use_pattern=True, use_group=False)
                v = lib.host_vbr_fill(N, N, rowptr, colind, None, g, 64, 64, force_fixed_size=False, pattern_only=True)
                blocks = len(v["jab"])
                dens = blocks / (v["block_rows"] * ((N + 63) // 64))
                for n in [int(x) for x in args.ns.split(",")]:
                    Bm = synth.seeded_B(N, n, seed=2)
                    h = sparta_b200.Handle.from_vbr(N, N, 64, v["row_part"], v["nzcount"], v["jab"], v["mab"], n_hint=n,
                                                    wide_tiles=args.wide_tiles)
                    h.set_B(Bm, N, n)
                    for _ in range(3):
                        h.run()
                    ms = float(np.median([h.run() for _ in range(10)]))
                    chk = bench.spot_check(h, v, 0, v["block_rows"], n, "bf16", Bm, None, dev, 0, n_check=16)
                    st = h.stats()
                    h.close()
                    rec = {"rows": order, "element_density": density, "tau": tau, "n": n, "nz_blocks": blocks,
                           "block_density": dens, "ms": ms, "tflops": 2.0 * v["nztot"] * n / ms / 1e9,
                           "max_rel_err": chk["max_rel_err"], "ok": chk["ok"], "team": st["team"],
                           "split_pieces": st["split_pieces"], "wide_tiles": st["wide_tiles"]}
                    results.append(rec)
                    print(f"density {density:.0e} rows {order:22s} tau {tau:.1f} n {n:5d}: {blocks:6d} blocks "
                          f"({100 * dens:5.1f} % of the grid)  {ms * 1e3:8.1f} us  {rec['tflops']:7.1f} TFLOP/s  "
                          f"err {chk['max_rel_err']:.1e} ok={chk['ok']}", flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
