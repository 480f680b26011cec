#!/bin/bash
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
for mode in 2 1; do
timeout 600 python scripts/trace_run.py rmat12_a5 0 cta_pair=$mode > gpurun_out/trace_small_$mode.txt 2>&1
echo "=== mode $mode"; head -40 gpurun_out/trace_small_$mode.txt | cut -c1-200
done
