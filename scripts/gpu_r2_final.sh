#!/bin/bash
# Round-2 closing run on one GPU: parity suite, both bench arms as the driver runs them, smoke(), config #5 sweep,
# the other entry points.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_n1.err | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err; echo "reference arm rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "roofline", {k: d["roofline"][k] for k in ("frac", "traffic", "pipe_tensor_active_pct", "hbm_frac_of_measured")})
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "same_result")}, d["e2e"].get("vbr_arrays", {}).get("ms_per_step_this_rank"))
print("cpu", d["cpu_baseline"])
print("check", d["check"], "launches", d["gpu_launches"], "clocks", d["clocks"])
r = json.loads(open("gpurun_out/r2_bench_reference_n1.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"])
PY
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --workload er14_fixed --no-cpu-baseline > gpurun_out/r2_bench_er14.json 2>/dev/null
timeout 900 python bench.py --workload rmat16_a4 --no-cpu-baseline > gpurun_out/r2_bench_rmat16_a4.json 2>/dev/null
timeout 900 python bench.py --workload rmat16_a5 --weighted --no-cpu-baseline > gpurun_out/r2_bench_a5_weighted.json 2>/dev/null
timeout 900 python bench.py --workload rmat16_a5 --precision tf32 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_a5_tf32.json 2>/dev/null
for f in er14 rmat16_a4 a5_weighted a5_tf32; do python - gpurun_out/r2_bench_$f.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d.get("e2e") or {}
print(sys.argv[1], d["dtype"], "%.4f ms %.1f TFLOP/s frac %.3f" % (d["ms_per_step"], d["value"], d["roofline"]["frac"]), "check", d["check"]["ok"], d["check"]["max_rel_err"], d["check"]["max_rel_err_vs_rounded_operands"], "e2e ms", e.get("ms_per_step"))
PY
done
timeout 900 python scripts/other_paths_time.py > gpurun_out/r2_other_entry_points.txt 2>&1; tail -6 gpurun_out/r2_other_entry_points.txt
timeout 1200 python scripts/config5_sweep.py --out gpurun_out/r2_config5_sweep.json > gpurun_out/r2_config5_sweep.txt 2>&1; tail -4 gpurun_out/r2_config5_sweep.txt
SKIP_TESTS=1 WORKLOADS="rmat16_a5:bf16" bash scripts/gpu_r2_ab.sh "--l2-slab-mb 80" "--l2-slab-mb 40" ""
