#!/bin/bash
# bash scripts/gpu_r2_multi.sh N "<workload> <extra bench args>" ...   (torchrun, one rank per GPU)
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ -n "$RUN_TESTS" ]; then timeout 600 python -m pytest tests/test_spmm_gpu.py -m gpu -q -k "multi_gpu" 2>&1 | tail -3; fi
i=0
for spec in "$@"; do
  set -- $spec; wl=$1; shift
  out=gpurun_out/r2_${wl}_n${N}_$i
  SPARTA_TIMING=$TIMING timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+i)) \
    bench.py --gpus $N --workload $wl "$@" > $out.json 2> $out.err; echo "bench $wl N=$N rc=$?"
  grep -E "sparta_csr|Error|error" $out.err | tail -4 | cut -c1-300
  python - $out.json <<'PY'
import json, sys
lines = [l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")]
if not lines:
    print("NO JSON LINE"); sys.exit(0)
d = json.loads(lines[-1])
e = d.get("e2e") or {}
print(f"{d['config']['workload']} N={d['n_gpus']} value {d['value']:.1f} TFLOP/s {d['ms_per_step']:.4f} ms frac {d['roofline']['frac']:.3f} hbm_frac {d['roofline']['hbm_frac_of_measured']:.3f} check {d['check']}")
print("  e2e", {k: e.get(k) for k in ("value", "ms_per_step", "ms_each_step_per_rank", "same_result")})
print("  setup", {k: d['setup'][k] for k in ("b_broadcast_s", "grid", "items", "gather_rows", "gather_nnz", "shard_block_rows")})
PY
  i=$((i+1))
done
