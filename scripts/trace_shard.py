"""Where does an eighth shard's launch time go?  Per-launch time of one shard of the bench workload next to
the timeline of several workers (first chunk issued, last drain done, items), plus the launch time of a
trivially small handle (fixed cost of a launch).  Diagnostic; needs a GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sparta_b200  # noqa: E402
from sparta_b200 import synth  # noqa: E402


def main():
    import torch
    wl = bench.WORKLOADS["rmat16_a5"]
    N, rowptr, colind = bench.make_matrix(wl)
    g = bench.make_grouping(wl, N, rowptr, colind)
    v = bench.build_vbr(wl, N, rowptr, colind, g)
    n = wl["n"]
    Bd = torch.from_numpy(synth.seeded_B(v["cols"], n, seed=2)).cuda()
    world = 8
    cuts = sparta_b200.partition_block_rows_modelled(v["rows"], v["cols"], 64, v["row_part"], v["nzcount"], v["jab"], n, world)
    for r in (0, 4, 7):
        for split in (0, 1):
            h = sparta_b200.Handle.from_vbr(v["rows"], v["cols"], 64, v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                            block_row_begin=int(cuts[r]), block_row_end=int(cuts[r + 1]), split_k=split)
            h.set_B_device(Bd.data_ptr(), v["cols"], n)
            stream = torch.cuda.ExternalStream(h.stream)
            for _ in range(3):
                h.run_async()
            h.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(20):
                h.run_async()
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / 20
            one = h.run()
            st = h.stats()
            print(f"shard {r} split_k={split}: {ms:.4f} ms per launch back to back, {one:.4f} ms alone (events around one launch); "
                  f"items {st['items']} grid {st['grid']} split pieces {st['split_pieces']} zero tiles {st['zero_tiles']} "
                  f"model {st['sched_max_cycles'] / 1e3:.0f} kcycles")
            spans = []
            for w in (0, 17, 35, 53, 71):
                if w * 2 >= st["grid"]:
                    continue
                rec = h.run_traced(w, 4096).astype(np.int64)
                prod, mma, epi = rec[0], rec[1], rec[2]
                nch = int((prod[0, :, 1] > 0).sum())
                nit = int((epi[0, :, 1] > 0).sum())
                if nch == 0 or nit == 0:
                    continue
                t0 = prod[0, 0, 0]
                last = epi[0, nit - 1, 1]
                drains = (epi[0, :nit, 1] - epi[0, :nit, 0])
                spans.append((w, nch, nit, int(last - t0), int(drains.sum()), int(mma[1, nch - 1, 1] - t0)))
            for w, nch, nit, span, dr, mm in spans:
                print(f"   worker {w}: {nch} chunks, {nit} items, first issue -> last drain {span} cycles ({span / 1.965e6:.4f} ms), "
                      f"of which drains {dr}, last MMA issued at {mm}")
            h.close()
    # fixed cost of a launch: a handle with almost nothing to do
    from tests.util import random_vbr
    rng = np.random.default_rng(0)
    tv = random_vbr(rng, 8, 512, 64, [64] * 8, 0.5, values="int")
    h = sparta_b200.Handle.from_vbr(tv["rows"], 512, 64, tv["row_part"], tv["nzcount"], tv["jab"], tv["mab"])
    h.set_B(rng.integers(-2, 3, size=(256, 512)).astype(np.float32), 512, 256)
    stream = torch.cuda.ExternalStream(h.stream)
    for _ in range(5):
        h.run_async()
    h.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(200):
        h.run_async()
    e1.record(stream)
    e1.synchronize()
    print(f"tiny handle (8 blocks): {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per launch back to back, {h.run() * 1e3:.1f} us alone, grid {h.stats()['grid']}")
    h.close()


if __name__ == "__main__":
    main()
