#!/bin/bash
set -x
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for ps in 4 6; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --panel-stages $ps > gpurun_out/bench_ps$ps.json 2> gpurun_out/bench_ps$ps.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_ps$ps.json'));print('PAIR ps$ps', d['value'], d['ms_per_step'], d['check']['ok'], d['clocks'])"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --cta-pair 1 > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err
python -c "import json;d=json.load(open('gpurun_out/bench_single.json'));print('SINGLE', d['value'], d['ms_per_step'], d['check']['ok'])"
timeout 600 python scripts/trace_run.py rmat16_a5 0 > gpurun_out/trace_pair.txt 2>&1
cat gpurun_out/trace_pair.txt | head -45
