#!/bin/bash
# round-1k GPU call: full parity suite (row ops, sum_into), default bench, launch list
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
echo "=== smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log | cut -c1-160
echo "=== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== bench default"; timeout 400 python bench.py --detail > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; grep "ms x" gpurun_out/bench_default.err; cut -c1-400 gpurun_out/bench_default.json
