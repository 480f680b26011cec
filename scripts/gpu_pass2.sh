#!/bin/bash
# pair kernel + L2-aware order: parity, then bench in both modes
set -x
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for mode in 2 1; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --cta-pair $mode > gpurun_out/bench_pair$mode.json 2> gpurun_out/bench_pair$mode.err; echo "bench rc=$?"
  cat gpurun_out/bench_pair$mode.json; tail -3 gpurun_out/bench_pair$mode.err
done
