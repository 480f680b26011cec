#!/bin/bash
# Round 2, first GPU call: multicast microbenchmark, baselines of the two weak workloads, ncu capture of rmat16_a4.
export SPARTA_BENCH_CACHE=cache
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 120 scripts/microbench/mcast_rate > gpurun_out/r2_mcast_rate.txt 2>&1; tail -30 gpurun_out/r2_mcast_rate.txt
timeout 600 python bench.py --workload rmat16_a4 --precision tf32 --no-e2e --no-cpu-baseline > gpurun_out/r2_a4_tf32_before.json 2> gpurun_out/r2_a4_tf32_before.err; tail -c 600 gpurun_out/r2_a4_tf32_before.json
timeout 300 python bench.py --workload er14_fixed --no-e2e --no-cpu-baseline > gpurun_out/r2_er14_before.json 2> gpurun_out/r2_er14_before.err; tail -c 400 gpurun_out/r2_er14_before.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_vbr -s 3 -c 1 -f -o gpurun_out/r2_prof_a4_before \
  python bench.py --workload rmat16_a4 --precision tf32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_a4.log 2>&1
ncu -i gpurun_out/r2_prof_a4_before.ncu-rep --page raw --csv > gpurun_out/r2_prof_a4_before_raw.csv 2>/dev/null
ls -la gpurun_out/r2_prof_a4_before.ncu-rep
