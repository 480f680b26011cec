#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/trace_run.py er14_fixed 0 > gpurun_out/r2_trace_er14.txt 2>&1; head -16 gpurun_out/r2_trace_er14.txt | cut -c1-220
SKIP_TESTS=1 WORKLOADS="er14_fixed:bf16" bash scripts/gpu_r2_ab.sh "--steps 50" "--steps 50 --acc-cols 256" "--steps 50 --split-k 2" "--steps 50 --acc-cols 256 --split-k 2" "--steps 50 --cta-pair 1" "--steps 50 --panel-stages 4 --pipeline 1"
timeout 900 python scripts/shard_scaling.py --partition model --split 0 --worlds 1,2,4,8 --out gpurun_out/r2_shard_scaling.json 2>&1 | tail -5
