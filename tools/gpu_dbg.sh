#!/bin/bash
# timing experiments: which part of the linear-system tile kernels costs what
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-dbg}; shift || true
i=0
for v in "$@"; do
  name=${TAG}_v$i
  echo "=== [$v]"
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --detail --sst --no-cpu-baseline > gpurun_out/$name.json 2> gpurun_out/$name.detail.txt
  echo "# variant: [$v]" >> gpurun_out/$name.detail.txt
  grep -E "ms x" gpurun_out/$name.detail.txt | grep -E "momentum|continuity|scalar"
  i=$((i+1))
done
