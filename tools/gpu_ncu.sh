#!/bin/bash
# ncu --set full of the linear-system tile kernels (one launch each) + launch list
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r02n}
CMD="python bench.py --steps 2 --warmup 3 --sst --no-cpu-baseline --sustain-s 0.01"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ls_tile_kernel -s 12 -c 4 \
  -o gpurun_out/${TAG}_ls_tile -f $CMD > gpurun_out/${TAG}_ncu_ls.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_ls.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv \
  --log-file gpurun_out/${TAG}_launches.csv $CMD > /dev/null 2>&1
tail -5 gpurun_out/${TAG}_launches.csv | cut -c1-200
ls -la gpurun_out/${TAG}_*
