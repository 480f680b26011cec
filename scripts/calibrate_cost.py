#!/usr/bin/env python
"""Calibration of the scheduler's cost model (csrc/schedule.cpp) against worker timelines.

For shards of the bench workload, whole-unit plans (split_k = 1): traces several workers with
sparta_run_traced and regresses the measured per-item period (accumulator drained -> next
accumulator drained) on the item's features: chunks, tensor cycles, bytes staged per CTA,
accumulator columns.  Prints the fit and the residuals by item weight class.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sparta_b200  # noqa: E402
from sparta_b200 import lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="rmat16_a5")
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--ranks", default="0,3,7")
    ap.add_argument("--workers", default="0,9,18,27,36")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    wl = bench.WORKLOADS[args.workload]
    N, rowptr, colind = bench.make_matrix(wl)
    grouping = bench.make_grouping(wl, N, rowptr, colind)
    v = bench.build_vbr(wl, N, rowptr, colind, grouping)
    n = wl["n"]
    from sparta_b200 import synth
    Bd = torch.from_numpy(synth.seeded_B(v["cols"], n, seed=2)).cuda()
    cuts = sparta_b200.partition_block_rows(v["row_part"], v["nzcount"], args.world)
    rows = []
    for r in [int(x) for x in args.ranks.split(",")]:
        lo, hi = int(cuts[r]), int(cuts[r + 1])
        plan = sparta_b200.vbr_plan(v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"], v["jab"], n,
                                    block_row_begin=lo, block_row_end=hi, split_k=1)
        h = sparta_b200.Handle.from_vbr(v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"], v["jab"],
                                        v["mab"], block_row_begin=lo, block_row_end=hi, split_k=1)
        h.set_B_device(Bd.data_ptr(), v["cols"], n)
        for _ in range(3):
            h.run()
        sr, ch = plan["srows"], plan["chunks"]
        for w in [int(x) for x in args.workers.split(",")]:
            if w >= len(plan["cta_ptr"]) - 1:
                continue
            rec = h.run_traced(worker=w, capacity=8192).astype(np.int64)
            epi = rec[2, 0]
            ids = plan["cta_items"][plan["cta_ptr"][w]:plan["cta_ptr"][w + 1]]
            for idx, it in enumerate(plan["items"][ids]):
                if idx == 0 or epi[idx, 1] == 0:
                    continue
                s = sr[it["srow"]]
                cnt = int(it["count"]) & L.ITEM_COUNT_MASK
                cs = ch[s["chunk_begin"] + it["chunk_off"]: s["chunk_begin"] + it["chunk_off"] + cnt]
                rows_present = cs["a_bytes"] / 128.0
                rows.append({"rank": r, "worker": w, "period": int(epi[idx, 1] - epi[idx - 1, 1]),
                             "epilogue": int(epi[idx, 1] - epi[idx, 0]), "chunks": cnt,
                             "tensor": float((cs["ksteps"] * rows_present * 0.5).sum()),
                             "bytes": float((16384 + cs["a_bytes"] / 2).sum()), "cols": int(s["n_cols"])})
        h.close()
    X = np.array([[1.0, d["chunks"], d["tensor"], d["bytes"], d["cols"]] for d in rows])
    y = np.array([d["period"] for d in rows], dtype=np.float64)
    coef, *_ = np.linalg.lstsq(X, y, rcond=None)
    print("items", len(rows))
    print("period ~ %.0f + %.1f*chunks + %.3f*tensor_cycles + %.5f*bytes_per_cta + %.1f*acc_cols" % tuple(coef))
    pred = X @ coef
    print("fit: median |err|/period = %.3f" % float(np.median(np.abs(pred - y) / y)))
    e = np.array([d["epilogue"] for d in rows], dtype=np.float64)
    c = np.array([d["cols"] for d in rows], dtype=np.float64)
    ce, *_ = np.linalg.lstsq(np.stack([np.ones_like(c), c], 1), e, rcond=None)
    print("epilogue ~ %.0f + %.1f*acc_cols  (median %.0f cycles at %d columns)" % (ce[0], ce[1], float(np.median(e)), int(np.median(c))))
    # per weight class: mean bytes per chunk
    bpc = X[:, 3] / np.maximum(X[:, 1], 1)
    for lo_, hi_ in ((0, 20000), (20000, 24000), (24000, 28000), (28000, 1e9)):
        m = (bpc >= lo_) & (bpc < hi_) & (X[:, 1] > 0)
        if m.any():
            print("  bytes/chunk/CTA in [%d, %d): %d items, period/chunk measured %.0f, tensor/chunk %.0f, chunks/item %.0f"
                  % (lo_, min(hi_, 99999), int(m.sum()), float((y[m] / X[m, 1]).mean()), float((X[m, 2] / X[m, 1]).mean()),
                     float(X[m, 1].mean())))
    z = X[:, 1] == 0
    if z.any():
        print("  zero-chunk items: %d, period %.0f" % (int(z.sum()), float(y[z].mean())))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"coef": coef.tolist(), "epilogue_coef": ce.tolist(), "items": rows}, f)


if __name__ == "__main__":
    main()
