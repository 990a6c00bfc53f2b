#!/bin/bash
# ncu --set full of the monolithic momentum tile kernel (one launch) and of a
# gradient kernel, exported as a small CSV of the metrics DESIGN.md quotes
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r02zn}
CMD="python bench.py --steps 2 --warmup 3 --detail --no-cpu-baseline --sustain-s 0.01"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:MomentumMonoP -s 1 -c 1 -o gpurun_out/${TAG}_mono -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__grid_size"
ncu -i gpurun_out/${TAG}_mono.ncu-rep --page raw --csv --metrics $M > gpurun_out/${TAG}_mono_summary.csv 2> /dev/null
cut -c1-600 gpurun_out/${TAG}_mono_summary.csv | tail -3
ls -la gpurun_out/${TAG}_*
