#!/bin/bash
# CSR kernel, bounded tf32 chains, pipelined one-shot call.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
SPARTA_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -E "^sparta" gpurun_out/bench.err | tail -4
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("bf16", d["value"], d["ms_per_step"], d["e2e"], d["check"])
PY
for mc in 0 -1 1024; do
timeout 900 python bench.py --precision tf32 --max-chain $mc --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_tf32_$mc.json 2> gpurun_out/bench_tf32_$mc.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_tf32_$mc.json").read().strip().splitlines()[-1])
print("tf32 max_chain=$mc", d["value"], d["ms_per_step"], d["check"], d["setup"]["items"])
PY
done
SPARTA_TIMING=1 timeout 600 python bench.py --workload er14_fixed --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_er14.json 2> gpurun_out/bench_er14.err
grep -E "^sparta" gpurun_out/bench_er14.err | tail -2
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_er14.json").read().strip().splitlines()[-1])
print("er14", d["value"], d["ms_per_step"], d["e2e"])
PY
timeout 600 python scripts/csr_time.py 2>&1 | tail -4
