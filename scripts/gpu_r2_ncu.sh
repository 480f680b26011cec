#!/bin/bash
# Round 2 ncu passes (B200_PROFILING.md): launch list of the default bench + one --set full capture of the dominant
# kernel of each workload (and of the gather kernel on the variable-height one).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e-vbr > gpurun_out/r2_ncu_list.log 2>&1
cap() {  # name workload precision kernel-regex extra
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$4 -s ${6:-3} -c 1 -f -o gpurun_out/r2_prof_$1 \
    python bench.py --workload $2 --precision $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline $5 > gpurun_out/r2_ncu_$1.log 2>&1
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_prof_$1_raw.csv 2>/dev/null
  ls -la gpurun_out/r2_prof_$1.ncu-rep | awk '{print $5, $9}'
}
cap a5 rmat16_a5 bf16 spmm_vbr ""
cap er14 er14_fixed bf16 spmm_vbr ""
cap a4_tc rmat16_a4 tf32 spmm_vbr ""
cap a4_gather rmat16_a4 tf32 spmm_csr ""
cap a4_tc_only rmat16_a4 tf32 spmm_vbr "--gather-max-height -1"
