#!/bin/bash
# Split (stream-K) assignment: parity, shard-scaling emulation on one GPU, bench line.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmm_gpu.py -m gpu -x -q -k "split or accumulate or shards" > gpurun_out/pytest_split.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_split.log
tail -5 gpurun_out/pytest_split.log
timeout 1200 python scripts/shard_scaling.py --out gpurun_out/shard_scaling.json 2>&1 | tee gpurun_out/shard_scaling.log | tail -12
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("bf16", d["value"], d["ms_per_step"], d["check"]["max_rel_err"], d["setup"])
PY
