#!/bin/bash
# Config #4 at full size on ONE GPU (does it run, how long), the weighted config #3, the new integration tests.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_integration_gpu.py tests/test_csr_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --workload rmat16_a5 --weighted --no-cpu-baseline > gpurun_out/r2_a5_weighted.json 2> gpurun_out/r2_a5_weighted.err; tail -c 1500 gpurun_out/r2_a5_weighted.json; tail -3 gpurun_out/r2_a5_weighted.err
timeout 1200 python bench.py --workload rmat18_a4 --steps 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_c4_n1.json 2> gpurun_out/r2_c4_n1.err; tail -c 1500 gpurun_out/r2_c4_n1.json; tail -5 gpurun_out/r2_c4_n1.err
