"""Per-item summary of a worker timeline (gpurun_out/trace_*.npy) against the host plan."""
import sys
import numpy as np

rec = np.load(sys.argv[1]).astype(np.int64)
plan = np.load(sys.argv[2], allow_pickle=True)
worker = int(sys.argv[3]) if len(sys.argv) > 3 else 0
nshare = int(sys.argv[4]) if len(sys.argv) > 4 else 2
prod, mma, epi, accw = rec
items = plan["items"][plan["cta_items"][plan["cta_ptr"][worker]:plan["cta_ptr"][worker + 1]]]
sr, ch = plan["srows"], plan["chunks"]
pos = 0
t0 = prod[0, 0, 0]
ntr = int((prod[0, :, 1] > 0).sum())
print("item srow j0 nchunk avgA_KB/CTA period issue idle latency ideal/chunk  start_ms  epi_cycles")
tot_ideal = tot_span = 0
for idx, it in enumerate(items):
    s = sr[it["srow"]]
    n = int(s["chunk_count"])
    if n == 0:
        continue
    if pos + n > ntr:
        break
    cs = ch[s["chunk_begin"]:s["chunk_begin"] + n]
    sl = slice(pos, pos + n)
    span = mma[1, pos + n - 1, 1] - mma[1, pos, 0]
    ideal = (cs["a_bytes"] / 128 / 2 * cs["ksteps"]).sum()
    tot_ideal += ideal
    tot_span += span
    idle = (mma[1, pos + 1:pos + n, 0] - mma[1, pos:pos + n - 1, 1]).mean() if n > 1 else 0
    print(it["srow"], it["j0"], n, round(cs["a_bytes"].mean() / nshare / 1024, 1), round(span / n),
          round((mma[1, sl, 1] - mma[1, sl, 0]).mean()), round(idle), round((mma[0, sl, 0] - prod[0, sl, 1]).mean()),
          round(ideal / n), round((prod[0, pos, 0] - t0) / 1.965e6, 3), epi[0, idx, 1] - epi[0, idx, 0])
    pos += n
print("traced chunks", pos, "sum ideal tensor cycles", tot_ideal, "sum item spans", tot_span, "ratio", tot_ideal / tot_span)
