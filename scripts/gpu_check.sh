#!/bin/bash
# Quick check of a scheduler change: parity suite, shard scaling, bench line.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python scripts/shard_scaling.py --partition model --split 0 --out gpurun_out/ss_check.json 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], d["check"]["ok"], "grid", d["setup"]["grid"])
PY
