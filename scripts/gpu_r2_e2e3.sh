#!/bin/bash
# Pinned staging of the library's own uploads: parity of everything that uploads, then the e2e phase breakdown
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SPARTA_TIMING=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; grep -E "^sparta" gpurun_out/r2p_bench_n1.err | grep -v "schedule:" | tail -24
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2p_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "check", d["check"]["ok"], "wide", d["setup"].get("wide_tiles"))
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "ms_each_step_per_rank", "same_result")})
print("e2e vbr", d["e2e"].get("vbr_arrays"))
PY
timeout 900 python bench.py --no-cpu-baseline --workload er14_fixed > gpurun_out/r2p_bench_er14.json 2>/dev/null
timeout 900 python bench.py --no-cpu-baseline --workload rmat16_a4 > gpurun_out/r2p_bench_a4.json 2>/dev/null
python - <<'PY'
import json
for f in ("er14", "a4"):
    d = json.loads(open(f"gpurun_out/r2p_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "ms", d["ms_per_step"], d["value"], "check", d["check"]["ok"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["ms_each_step_per_rank"], d["e2e"]["same_result"])
PY
