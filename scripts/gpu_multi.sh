#!/bin/bash
# Real multi-GPU run of the bench (one rank per GPU, NCCL): $1 = number of GPUs.
N=${1:-2}
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/gpus_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
lines = [l for l in open("gpurun_out/bench_n$N.json").read().strip().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "check", d["check"], "setup", d["setup"])
PY
[ "$N" -le 2 ] && timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N rc=$?"
[ "$N" -le 2 ] && tail -2 gpurun_out/bench_ref_n$N.json | cut -c1-300
