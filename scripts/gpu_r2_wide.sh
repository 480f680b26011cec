#!/bin/bash
# Wide items (several column tiles per work item) on one box: parity first, then the A/B over wide_tiles
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_spmm_gpu.py -m gpu -x -q -k "wide" 2>&1 | tail -15
timeout 1500 python -m pytest tests/test_spmm_gpu.py -m gpu -x -q -k "not wide" 2>&1 | tail -4
SKIP_TESTS=1 WORKLOADS="er14_fixed:bf16" bash scripts/gpu_r2_ab.sh "--wide-tiles 1" "--wide-tiles 2" "--wide-tiles 4" "" "--wide-tiles 4 --cta-pair 1" "--wide-tiles 2 --cta-pair 1"
SKIP_TESTS=1 WORKLOADS="rmat16_a5:bf16 rmat16_a4:tf32" bash scripts/gpu_r2_ab.sh "" "--wide-tiles 2"
