#!/bin/bash
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
bash scripts/gpu_sweep.sh "--cta-pair 1" "--cta-pair 1 --acc-cols 256" "--cta-pair 1 --acc-cols 256 --panel-stages 6" "--cta-pair 2 --panel-stages 8" "--cta-pair 2 --acc-cols 256 --panel-stages 8" "--cta-pair 2 --acc-cols 256 --panel-stages 6" "--cta-pair 1 --panel-stages 3" "--cta-pair 1 --panel-stages 2"
timeout 600 python scripts/trace_run.py rmat16_a5 0 cta_pair=1 > gpurun_out/trace_single.txt 2>&1
head -12 gpurun_out/trace_single.txt | cut -c1-200
cp gpurun_out/trace_rmat16_a5_w0.npy gpurun_out/trace_single_w0.npy
