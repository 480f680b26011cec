#!/bin/bash
# Wide items as the default: whole GPU suite, the bench with the e2e phase breakdown, shard scaling on one GPU
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SPARTA_TIMING=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; grep -E "^sparta" gpurun_out/r2w_bench_n1.err | tail -8
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "check", d["check"]["ok"], "wide", d["setup"].get("wide_tiles"))
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "same_result")})
PY
timeout 900 python scripts/shard_scaling.py --partition model --split 0 --out gpurun_out/r2w_shard_scaling.json 2>&1 | grep -v "^\[bench\]"
timeout 600 python scripts/shard_scaling.py --partition model --split 0 --worlds 1,8 --opts wide_tiles=1 2>&1 | grep -v "^\[bench\]"
timeout 600 python scripts/shard_scaling.py --partition model --split 0 --worlds 1,8 --opts l2_slab_mb=300 2>&1 | grep -v "^\[bench\]"
