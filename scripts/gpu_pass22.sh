#!/bin/bash
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for prec in tf32 bf16; do
timeout 900 python bench.py --workload rmat16_a4 --precision $prec --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_a4_$prec.json 2> gpurun_out/bench_a4_$prec.err; echo "a4 $prec rc=$?"
done
timeout 900 python bench.py --precision tf32 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "tf32 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bf16 rc=$?"
python - <<'PY'
import json
for f in ("bench_a4_tf32", "bench_a4_bf16", "bench_tf32", "bench"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d["dtype"], round(d["value"], 1), "TFLOP/s", round(d["ms_per_step"], 4), "ms", "e2e", d["e2e"] and round(d["e2e"]["value"], 2), "err", d["check"]["max_rel_err"], d["check"]["ok"], d["setup"]["items"], d["setup"]["a_upload_pack_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
