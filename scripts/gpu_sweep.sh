#!/bin/bash
# usage: gpu_sweep.sh "<bench args 1>" "<bench args 2>" ...
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
i=0
for args in "$@"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $args > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err || tail -5 gpurun_out/sweep_$i.err
  python -c "import json,sys;d=json.load(open('gpurun_out/sweep_$i.json'));print('SWEEP [$args]', round(d['value'],1), 'TF/s', round(d['ms_per_step'],3), 'ms ok=%s' % d['check']['ok'], 'team', d['setup'].get('team'), 'W', d['clocks']['power_w_max'])"
  i=$((i+1))
done
