#!/bin/bash
# round-1k GPU call: full parity suite (row ops, node kernels, sum_into) + default bench
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
echo "=== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
