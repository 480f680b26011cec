#!/usr/bin/env python
"""Strong-scaling emulation on ONE GPU: every rank's shard of the bench workload is run by
itself and timed with CUDA events; the N-GPU step time is the slowest shard (there is no
collective inside the multiply), so efficiency(N) = T(1) / (N * max_r T_r(N)).

    python scripts/shard_scaling.py [--workload rmat16_a5] [--worlds 1,2,4,8] [--split 0,1]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="rmat16_a5")
    ap.add_argument("--worlds", default="1,2,4,8")
    ap.add_argument("--split", default="1,0", help="split_k settings to compare (1 never, 0 auto, 2 always)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default=None)
    ap.add_argument("--opts", default="", help="extra handle options, e.g. num_ctas=144,l2_slab_mb=160")
    ap.add_argument("--partition", default="area", choices=["area", "model"])
    ap.add_argument("--feedback", type=int, default=0, help="re-cut this many times with measured / modelled shard times")
    args = ap.parse_args()
    import torch
    wl = bench.WORKLOADS[args.workload]
    N, rowptr, colind = bench.make_matrix(wl)
    grouping = bench.make_grouping(wl, N, rowptr, colind)
    v = bench.build_vbr(wl, N, rowptr, colind, grouping)
    n = wl["n"]
    from sparta_b200 import synth
    Bd = torch.from_numpy(synth.seeded_B(v["cols"], n, seed=2)).cuda()
    total_flops = 2.0 * v["nztot"] * n
    extra = {k: int(x) for k, x in (kv.split("=") for kv in args.opts.split(",") if kv)}
    results = []
    def measure(cuts, world, split):
        per_rank = []
        for r in range(world):
            h = sparta_b200.Handle.from_vbr(v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"], v["jab"],
                                            v["mab"], precision=args.precision, block_row_begin=int(cuts[r]),
                                            block_row_end=int(cuts[r + 1]), split_k=split, n_hint=n, **extra)
            h.set_B_device(Bd.data_ptr(), v["cols"], n)
            st = h.stats()
            stream = torch.cuda.ExternalStream(h.stream)
            for _ in range(3):
                h.run_async()
            h.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                h.run_async()
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            per_rank.append({"rank": r, "ms": ms, "nz_blocks": st["nz_blocks"], "items": st["items"],
                             "split_pieces": st["split_pieces"], "zero_tiles": st["zero_tiles"],
                             "model_imbalance": round(st["sched_imbalance"], 3),
                             "model_cycles": st["sched_max_cycles"], "chunks": st["chunks"],
                             "super_rows": st["super_rows"],
                             "tflops": 2.0 * st["nztot"] * n / ms / 1e9})
            h.close()
        return per_rank

    for split in [int(x) for x in args.split.split(",")]:
        t1 = None
        for world in [int(x) for x in args.worlds.split(",")]:
            if args.partition == "model":
                cuts = sparta_b200.partition_block_rows_modelled(v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"],
                                                                 v["jab"], n, world, precision=args.precision,
                                                                 split_k=split, **extra)
            else:
                cuts = sparta_b200.partition_block_rows(v["row_part"], v["nzcount"], world)
            for attempt in range(args.feedback + 1 if world > 1 else 1):
                if attempt > 0:   # re-cut with measured / modelled shard times of the previous attempt
                    cuts = sparta_b200.partition_block_rows_measured(
                        v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"], v["jab"], n, world, cuts,
                        [p["ms"] for p in per_rank], [p["model_cycles"] for p in per_rank], precision=args.precision,
                        split_k=split, **extra)
                per_rank = measure(cuts, world, split)
                worst = max(p["ms"] for p in per_rank)
                if world == 1:
                    t1 = worst
                rec = {"split_k": split, "world": world, "feedback_round": attempt, "cuts": [int(c) for c in cuts],
                       "step_ms": worst, "tflops": total_flops / worst / 1e9,
                       "efficiency": (t1 / (world * worst)) if t1 else None, "per_rank": per_rank}
                results.append(rec)
                print(f"split_k={split} N={world}" + (f" feedback {attempt}" if attempt else "") +
                      f": step {worst:.3f} ms, {rec['tflops']:.0f} TFLOP/s aggregate, "
                      f"efficiency {rec['efficiency']:.3f}; per-rank ms " + " ".join(f"{p['ms']:.3f}" for p in per_rank),
                      flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
