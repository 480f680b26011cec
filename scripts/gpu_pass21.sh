#!/bin/bash
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spmm_gpu.py -m gpu -x -q -k "variable_height or unpermuted" 2>&1 | tail -3
for prec in tf32 bf16; do
timeout 900 python bench.py --workload rmat16_a4 --precision $prec --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_a4_$prec.json 2> gpurun_out/bench_a4_$prec.err; echo "a4 $prec rc=$?"
tail -2 gpurun_out/bench_a4_$prec.err
done
python - <<'PY'
import json
for f in ("bench_a4_tf32", "bench_a4_bf16"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d["dtype"], round(d["value"], 1), "TFLOP/s", round(d["ms_per_step"], 4), "ms", "hbm frac", round(d["roofline"]["hbm_frac_of_measured"], 3),
              "e2e", d["e2e"] and round(d["e2e"]["value"], 2), "err", d["check"], d["setup"])
    except Exception as e:
        print(f, "failed", e)
PY
