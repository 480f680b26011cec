"""Timeline of one worker of the SpMM kernel on the bench workload (diagnostic; needs a GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sparta_b200  # noqa: E402
from sparta_b200 import synth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "rmat16_a5"
    worker = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    opts = dict(kv.split("=") for kv in sys.argv[3:])
    opts = {k: int(v) for k, v in opts.items()}
    wl = bench.WORKLOADS[name]
    N, rowptr, colind = bench.make_matrix(wl)
    g = bench.make_grouping(wl, N, rowptr, colind)
    v = bench.build_vbr(wl, N, rowptr, colind, g)
    h = sparta_b200.Handle.from_vbr(v["rows"], v["cols"], wl["w"], v["row_part"], v["nzcount"], v["jab"], v["mab"],
                                    precision="bf16", **opts)
    h.set_B(synth.seeded_B(v["cols"], wl["n"], 2), v["cols"], wl["n"])
    for _ in range(3):
        h.run()
    rec = h.run_traced(worker, 16384).astype(np.int64)
    dt = h.run()
    print("launch ms", dt, "stats", h.stats())
    prod, mma, epi, accw = rec[0], rec[1], rec[2], rec[3]
    nch = int((prod[0, :, 1] > 0).sum())
    nit = int((epi[0, :, 1] > 0).sum())
    t0 = prod[0, 0, 0]
    print("chunks traced", nch, "items", nit)
    if nch == 0:
        return
    span = mma[1, nch - 1, 1] - t0
    print("span cycles", span, "per chunk", span / nch)
    p_wait = (prod[0, :nch, 1] - prod[0, :nch, 0])
    print("producer: time inside issue (incl. waits) mean", p_wait.mean(), "p50", np.median(p_wait), "p90", np.percentile(p_wait, 90))
    gap = np.diff(prod[0, :nch, 0])
    print("producer: entry-to-entry mean", gap.mean(), "p50", np.median(gap))
    lat = mma[0, :nch, 0] - prod[0, :nch, 1]
    print("own load latency (TMA issued -> seen full by MMA thread): mean", lat.mean(), "p50", np.median(lat), "p90", np.percentile(lat, 90), "min", lat.min())
    if rec[1][0, :nch, 1].any():
        pw = mma[0, :nch, 1] - mma[0, :nch, 0]
        print("extra wait for the peer's stage: mean", pw.mean(), "p50", np.median(pw), "p90", np.percentile(pw, 90))
        plat = mma[0, :nch, 1] - prod[1, :nch, 1]
        print("peer load latency (peer TMA issued -> leader sees relay): mean", plat.mean(), "p50", np.median(plat))
    iss = mma[1, :nch, 1] - mma[1, :nch, 0]
    print("MMA issue time per chunk: mean", iss.mean(), "p50", np.median(iss))
    idle = mma[0, 1:nch, 0] - mma[1, :nch - 1, 1]
    print("MMA thread idle before next full: mean", idle.mean(), "p50", np.median(idle), "frac>0", (idle > 0).mean())
    e = epi[0, :nit, 1] - epi[0, :nit, 0]
    print("epilogue per item cycles: mean", e.mean(), "max", e.max())
    aw = accw[0, :nit, 1] - accw[0, :nit, 0]
    print("MMA wait for accumulator per item: mean", aw.mean(), "max", aw.max())
    np.save(os.path.join(ROOT, "gpurun_out", f"trace_{name}_w{worker}.npy"), rec)
    # first 24 chunks raw
    for i in range(min(24, nch)):
        print(i, "prod", prod[0, i] - t0, "peerprod", prod[1, i] - t0, "mma own/peer/done", mma[0, i, 0] - t0, mma[0, i, 1] - t0, mma[1, i, 1] - t0)


if __name__ == "__main__":
    main()
