#!/bin/bash
# In-kernel zeroing of split tiles, wider teams, modelled partition: parity + scaling + bench.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/shard_scaling.py --partition model --split 0,1 --out gpurun_out/shard_scaling_model.json 2>&1 | tee gpurun_out/shard_scaling_model.log | tail -9
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("bf16", d["value"], d["ms_per_step"], d["check"]["max_rel_err"], d["e2e"]["value"], d["setup"])
PY
