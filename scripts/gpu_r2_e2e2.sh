#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SPARTA_TIMING=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; grep -E "sparta" gpurun_out/r2_bench_n1.err | tail -6
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "check", d["check"]["ok"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "same_result", "max_rel_diff_vs_resident_handle")})
print("e2e vbr arrays", d["e2e"].get("vbr_arrays"))
PY
