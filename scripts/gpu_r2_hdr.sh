#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SKIP_TESTS=1 bash scripts/gpu_r2_ab.sh "--steps 20" "--steps 20 --copy-warps 1" "--steps 20 --pipeline 1"
timeout 300 python scripts/trace_run.py er14_fixed 0 2>&1 | head -14 | cut -c1-200
SPARTA_TIMING=1 timeout 900 python bench.py --no-cpu-baseline --no-e2e-vbr --steps 5 2> gpurun_out/e2e_t.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e', d['e2e']['ms_per_step'], d['e2e']['same_result'], 'kernel', d['ms_per_step'])"
grep sparta_csr gpurun_out/e2e_t.err | tail -3
