#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
SKIP_TESTS=1 WORKLOADS="rmat16_a4:tf32" bash scripts/gpu_r2_ab.sh "--gather-max-height -1" "--gather-max-height 0" "--gather-max-height 16" "--gather-max-height 64"
timeout 1200 python bench.py --workload rmat18_a4 --steps 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_c4_n1_hybrid.json 2> gpurun_out/r2_c4_n1_hybrid.err; tail -c 1200 gpurun_out/r2_c4_n1_hybrid.json; tail -3 gpurun_out/r2_c4_n1_hybrid.err
