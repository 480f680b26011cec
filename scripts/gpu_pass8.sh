#!/bin/bash
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_sweep.sh "--cta-pair 2" "--cta-pair 2 --panel-stages 6" "--cta-pair 2 --panel-stages 8" "--cta-pair 1"
timeout 600 python scripts/trace_run.py rmat12_a5 0 cta_pair=2 > gpurun_out/trace_small_2.txt 2>&1
head -16 gpurun_out/trace_small_2.txt | cut -c1-200
timeout 600 python scripts/trace_run.py rmat16_a5 0 cta_pair=2 > gpurun_out/trace_pair.txt 2>&1
head -12 gpurun_out/trace_pair.txt | cut -c1-200
