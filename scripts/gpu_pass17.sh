#!/bin/bash
# Recalibrated cost model, modelled partition, group-width variants at 8 shards.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 900 python scripts/shard_scaling.py --partition model --split 0 --out gpurun_out/shard_scaling_model.json 2>&1 | tee gpurun_out/shard_scaling_model.log | tail -6
timeout 900 python scripts/shard_scaling.py --partition area --split 0 --worlds 1,2,4,8 --out gpurun_out/shard_scaling_area.json 2>&1 | tee gpurun_out/shard_scaling_area.log | tail -6
timeout 900 python scripts/shard_scaling.py --partition model --split 0 --worlds 1,4,8 --opts num_ctas=144,l2_slab_mb=160 --out gpurun_out/shard_scaling_t4.json 2>&1 | tee gpurun_out/shard_scaling_t4.log | tail -6
timeout 900 python scripts/shard_scaling.py --partition model --split 0 --worlds 1,8 --opts num_ctas=144,l2_slab_mb=300 --out gpurun_out/shard_scaling_t8.json 2>&1 | tee gpurun_out/shard_scaling_t8.log | tail -6
