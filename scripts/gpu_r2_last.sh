#!/bin/bash
# Last kernel change of the round (one-tile schedules back on their own instantiation): parity, the shapes it was
# meant to restore, then the ncu passes on the frozen sources
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SKIP_TESTS=1 WORKLOADS="rmat16_a5:tf32 rmat16_a4:tf32" bash scripts/gpu_r2_ab.sh ""
SKIP_TESTS=1 WORKLOADS="er14_fixed:bf16" bash scripts/gpu_r2_ab.sh "--wide-tiles 1" ""
SKIP_TESTS=1 WORKLOADS="rmat16_a5:bf16" bash scripts/gpu_r2_ab.sh "--wide-tiles 1" ""
timeout 900 python scripts/other_paths_time.py > gpurun_out/r2_other_entry_points.txt 2>&1; tail -5 gpurun_out/r2_other_entry_points.txt
timeout 1200 bash scripts/gpu_r2_ncu.sh
