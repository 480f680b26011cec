#!/bin/bash
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_sweep.sh "--cta-pair 2" "--cta-pair 2 --panel-stages 5" "--cta-pair 2 --panel-stages 6" "--cta-pair 1" "--cta-pair 2 --panel-stages 6 --l2-slab-mb 40" "--cta-pair 2 --panel-stages 6 --row-order 1"
timeout 600 python scripts/trace_run.py rmat16_a5 0 cta_pair=2 panel_stages=6 > gpurun_out/trace_pair.txt 2>&1
head -12 gpurun_out/trace_pair.txt | cut -c1-200
