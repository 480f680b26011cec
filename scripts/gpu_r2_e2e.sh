#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_spmm_gpu.py -m gpu -x -q -k "from_csr or short_block or one_shot" 2>&1 | tail -5
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "check", d["check"]["ok"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "same_result", "max_rel_diff_vs_resident_handle")})
print("e2e vbr arrays", d["e2e"].get("vbr_arrays"))
PY
SPARTA_TIMING=1 timeout 600 python bench.py --no-cpu-baseline --steps 3 --no-e2e-vbr 2>&1 | grep -E "sparta|create" | tail -8
SKIP_TESTS=1 WORKLOADS="rmat18_a4:tf32" bash scripts/gpu_r2_ab.sh "--steps 5 --gather-passes 1" "--steps 5" "--steps 5 --gather-passes 8"
