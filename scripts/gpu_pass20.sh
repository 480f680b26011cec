#!/bin/bash
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 900 python scripts/shard_scaling.py --partition model --split 0 --worlds 1,8 --opts l2_slab_mb=300 --out gpurun_out/ss_t8.json 2>&1 | tail -3
timeout 900 python scripts/shard_scaling.py --partition model --split 0 --worlds 1,8 --opts l2_slab_mb=80 --out gpurun_out/ss_t2.json 2>&1 | tail -3
timeout 900 python scripts/shard_scaling.py --partition model --split 0 --worlds 1,4,8 --opts acc_cols=256 --out gpurun_out/ss_acc256.json 2>&1 | tail -4
