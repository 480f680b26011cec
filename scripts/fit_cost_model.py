#!/usr/bin/env python
"""Fit of the scheduler's per-chunk / per-item cost terms to measured shard times.

Input: the JSON written by scripts/shard_scaling.py (per-shard kernel times on one GPU).  For each
shard run with a split plan (balanced by construction) the plan is rebuilt on the CPU and the
per-worker sums of {items, chunk visits, A rows staged} are taken for the worker with the largest
modelled load; the measured time is regressed on them:  cycles ~ a*items + b*chunks + c*rows.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparta_b200  # noqa: E402
from sparta_b200 import lib as L  # noqa: E402


def main():
    res = json.load(open(sys.argv[1]))
    idx = np.load(sys.argv[2])          # row_part / nzcount / jab of the workload
    rp, nz, jab = idx["row_part"], idx["nzcount"], idx["jab"]
    n, cols, w = int(sys.argv[3]) if len(sys.argv) > 3 else 2048, 65536, 64
    partition = sys.argv[4] if len(sys.argv) > 4 else "area"
    clock = 1.965e9
    X, y, tag = [], [], []
    for rec in res:
        world = rec["world"]
        if partition == "model":
            cuts = sparta_b200.partition_block_rows_modelled(int(rp[-1]), cols, w, rp, nz, jab, n, world, split_k=rec["split_k"])
        else:
            cuts = sparta_b200.partition_block_rows(rp, nz, world)
        for pr in rec["per_rank"]:
            r = pr["rank"]
            plan = L.vbr_plan(int(rp[-1]), cols, w, rp, nz, jab, n, block_row_begin=int(cuts[r]),
                              block_row_end=int(cuts[r + 1]), split_k=rec["split_k"])
            items, sr, ch = plan["items"], plan["srows"], plan["chunks"]
            rows_of_chunk = ch["a_bytes"] / 128.0
            csum = np.concatenate([[0.0], np.cumsum(rows_of_chunk)])
            feats = []
            for wk in range(len(plan["cta_ptr"]) - 1):
                ids = plan["cta_items"][plan["cta_ptr"][wk]:plan["cta_ptr"][wk + 1]]
                it = items[ids]
                cnt = (it["count"] & L.ITEM_COUNT_MASK).astype(np.int64)
                c0 = sr["chunk_begin"][it["srow"]] + it["chunk_off"]
                feats.append((len(ids), int(cnt.sum()), float((csum[c0 + cnt] - csum[c0]).sum())))
            feats = np.array(feats)
            # the worker that decides: largest under a neutral prior (chunks + rows/64)
            lead = feats[np.argmax(feats[:, 1] * 600 + feats[:, 2] * 1.7 + feats[:, 0] * 20000)]
            X.append(lead)
            y.append(pr["ms"] * 1e-3 * clock)
            tag.append((rec["split_k"], world, r))
    X, y = np.array(X), np.array(y)
    Xi = np.concatenate([np.ones((len(X), 1)), X], axis=1)
    ci, *_ = np.linalg.lstsq(Xi, y, rcond=None)
    print(f"with intercept: cycles ~ {ci[0]:.0f} + {ci[1]:.0f}*items + {ci[2]:.1f}*chunks + {ci[3]:.3f}*rows   "
          f"median |err| {np.median(np.abs(Xi @ ci - y) / y):.3f}, max {np.max(np.abs(Xi @ ci - y) / y):.3f}")
    for sel, name in ((np.array([t[0] == 0 for t in tag]), "split plans"), (np.ones(len(tag), bool), "all plans")):
        coef, *_ = np.linalg.lstsq(X[sel], y[sel], rcond=None)
        pred = X[sel] @ coef
        print(f"{name}: cycles ~ {coef[0]:.0f}*items + {coef[1]:.1f}*chunks + {coef[2]:.3f}*rows   "
              f"median |err| {np.median(np.abs(pred - y[sel]) / y[sel]):.3f}, max {np.max(np.abs(pred - y[sel]) / y[sel]):.3f}")
    for t, x, yy in zip(tag, X, y):
        print(t, x.tolist(), round(yy))


if __name__ == "__main__":
    main()
