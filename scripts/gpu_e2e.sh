#!/bin/bash
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
export SPARTA_TIMING=1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "bench rc=$?"
grep -E "sparta" gpurun_out/bench_e2e.err | tail -12
python -c "import json;d=json.load(open('gpurun_out/bench_e2e.json'));print(d['value'], d['e2e'])"
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
