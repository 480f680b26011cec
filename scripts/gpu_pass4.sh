#!/bin/bash
set -x
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_a.json'));print('PAIR', d['value'], d['ms_per_step'], d['check']['ok'])"
timeout 600 python scripts/trace_run.py rmat16_a5 0 > gpurun_out/trace_pair.txt 2>&1
cat gpurun_out/trace_pair.txt | head -60
timeout 600 python scripts/trace_run.py rmat16_a5 0 cta_pair=1 > gpurun_out/trace_single.txt 2>&1
cat gpurun_out/trace_single.txt | head -40
