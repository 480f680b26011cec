#!/bin/bash
# Programmatic dependent launch between back-to-back multiplies: modes 0 (off), 1 (epilogue waits), 2 (MMA waits too)
mkdir -p gpurun_out
for e in "SPARTA_PDL_MODE=2" "SPARTA_PDL_MODE=0" "SPARTA_PDL_MODE=1"; do
  echo "== $e"
  env $e timeout 600 python -m pytest tests/test_spmm_gpu.py -m gpu -x -q -k "back_to_back or split or accumulate" 2>&1 | tail -1
  for w in rmat16_a5:bf16 er14_fixed:bf16; do
    name=${w%%:*}; prec=${w##*:}
    env $e timeout 600 python bench.py --workload $name --precision $prec --no-e2e --no-cpu-baseline > gpurun_out/pdl.json 2>/dev/null
    python - "$name" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/pdl.json").read().strip().splitlines()[-1])
print(f"{sys.argv[1]:12s} {d['ms_per_step']:.4f} ms {d['value']:.1f} TFLOP/s check {d['check']['ok']}")
PY
  done
  env $e timeout 900 python scripts/shard_scaling.py --partition model --split 0 --worlds 1,2,4,8 2>&1 | grep -v "^\[bench\]"
done
