#!/bin/bash
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
SPARTA_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bf16 rc=$?"
grep -E "^sparta" gpurun_out/bench.err | tail -3
SPARTA_TIMING=1 SPARTA_DENSE_UPLOAD=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err; echo "dense rc=$?"
grep -E "^sparta" gpurun_out/bench_dense.err | tail -2
timeout 900 python bench.py --workload rmat16_a4 --precision tf32 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_a4_tf32.json 2> gpurun_out/bench_a4_tf32.err; echo "a4 rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_dense", "bench_a4_tf32"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d["dtype"], round(d["value"], 1), "TFLOP/s", round(d["ms_per_step"], 4), "ms", "e2e", d["e2e"] and (round(d["e2e"]["value"], 2), round(d["e2e"]["ms_per_step"], 1)), "err", d["check"]["max_rel_err"], d["check"]["ok"], d["setup"]["a_upload_pack_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
