#!/bin/bash
# Epilogue drain microbenchmark; empty super-rows skipped; un-permuted read-back; shard scaling.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 300 ./scripts/microbench/epilogue_rate 2>&1 | tee gpurun_out/epilogue_rate.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python scripts/shard_scaling.py --out gpurun_out/shard_scaling.json 2>&1 | tee gpurun_out/shard_scaling.log | tail -12
