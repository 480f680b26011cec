#!/bin/bash
# ncu launch list of the bench + one full capture of the SpMM kernel (the two passes of B200_PROFILING.md).
export SPARTA_BENCH_CACHE=/tmp/sparta_cache
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_vbr -s 3 -c 1 -f -o gpurun_out/prof_spmm \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
ncu -i gpurun_out/prof_spmm.ncu-rep --page raw --csv > gpurun_out/prof_spmm_raw.csv 2>/dev/null
ls -la gpurun_out/prof_spmm.ncu-rep gpurun_out/launches.csv
