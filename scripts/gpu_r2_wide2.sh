#!/bin/bash
# Wide items, second pass: which shapes gain from 2 tiles per item
mkdir -p gpurun_out
SKIP_TESTS=1 WORKLOADS="rmat16_a5:bf16" bash scripts/gpu_r2_ab.sh "--wide-tiles 1" "--wide-tiles 2" "--wide-tiles 4" "--wide-tiles 1 --weighted" "--wide-tiles 2 --weighted" "--wide-tiles 2 --max-chain -1 --precision tf32"
SKIP_TESTS=1 WORKLOADS="er14_fixed:bf16" bash scripts/gpu_r2_ab.sh "--wide-tiles 2 --split-k 2" "--wide-tiles 2 --panel-stages 4" "--wide-tiles 2 --copy-warps 1"
SKIP_TESTS=1 WORKLOADS="rmat16_a4:bf16" bash scripts/gpu_r2_ab.sh "--wide-tiles 1" "--wide-tiles 2"
for t in 1 2; do
  echo "config5 sweep wide_tiles=$t"
  timeout 600 python scripts/config5_sweep.py --densities 5e-4,8e-3 --taus 0.0,0.6 --ns 512,1024,4096 --wide-tiles $t 2>&1 | grep -v "^\[bench\]"
done
