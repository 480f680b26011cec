#!/bin/bash
# Pipelined epilogue drain: full parity suite, calibration at 2 and 8 shards, shard scaling.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/calibrate_cost.py --world 2 --ranks 0,1 --workers 0,7,15,22,30,36 --out gpurun_out/calibrate_w2.json 2>&1 | tee gpurun_out/calibrate_w2.log | tail -12
timeout 600 python scripts/calibrate_cost.py --world 8 --ranks 0,4,7 --out gpurun_out/calibrate_w8.json 2>&1 | tee gpurun_out/calibrate_w8.log | tail -12
timeout 1200 python scripts/shard_scaling.py --out gpurun_out/shard_scaling.json 2>&1 | tee gpurun_out/shard_scaling.log | tail -12
