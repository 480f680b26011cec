#!/bin/bash
# A/B on one box: bash scripts/gpu_r2_ab.sh "<bench opts A>" "<bench opts B>" ...   (an option string may start
# with LIB=<path> to select another build of the library)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests/test_spmm_gpu.py -m gpu -x -q 2>&1 | tail -3; fi
i=0
for o in "$@"; do
  libsel=""
  case "$o" in LIB=*) libsel="${o%% *}"; libsel="${libsel#LIB=}"; o="${o#* }";; esac
  for w in ${WORKLOADS:-"rmat16_a5:bf16" "er14_fixed:bf16" "rmat16_a4:tf32"}; do
    name=${w%%:*}; prec=${w##*:}
    SPARTA_B200_LIB=$libsel timeout 600 python bench.py --workload $name --precision $prec --no-e2e --no-cpu-baseline $o > gpurun_out/ab_${name}_$i.json 2> gpurun_out/ab_${name}_$i.err
    python - "$name" "$libsel $o" gpurun_out/ab_${name}_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:12s} [{sys.argv[2]}] {d['ms_per_step']:.4f} ms {d['value']:.1f} TFLOP/s frac {d['roofline']['frac']:.3f} check {d['check']['ok']} grid {d['setup']['grid']} items {d['setup']['items']}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
  done
  i=$((i+1))
done
