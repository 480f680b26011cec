#!/bin/bash
# Inverted product (-M 6 / -M 11) parity + cost-model calibration from worker timelines.
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmm_gpu.py tests/test_integration_gpu.py -m gpu -x -q -k "inverted or cli" > gpurun_out/pytest_ba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ba.log
tail -8 gpurun_out/pytest_ba.log
timeout 1200 python scripts/calibrate_cost.py --out gpurun_out/calibrate.json 2>&1 | tee gpurun_out/calibrate.log | tail -20
