#!/bin/bash
# Final profile refresh: launch list + full capture of the kernel as committed, bench lines.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_vbr -s 3 -c 1 -f -o gpurun_out/prof_spmm \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
ncu -i gpurun_out/prof_spmm.ncu-rep --page raw --csv > gpurun_out/prof_spmm_raw.csv 2>/dev/null
timeout 900 python bench.py --precision tf32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "tf32 rc=$?"
timeout 600 python bench.py --workload er14_fixed --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_er14.json 2> gpurun_out/bench_er14.err; echo "er14 rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_tf32", "bench_er14"):
    d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print(f, d["dtype"], round(d["value"], 1), "TFLOP/s", round(d["ms_per_step"], 4), "ms frac", round(d["roofline"]["frac"], 3), "e2e", d["e2e"] and (round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1)), d["check"]["max_rel_err"], d["clocks"])
PY
