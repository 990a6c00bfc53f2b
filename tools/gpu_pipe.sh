#!/bin/bash
# first runs of the warp-specialised pipelined kernel: parity, then A/B by tile size
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== pytest pipe"
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pipe_kernel" > gpurun_out/r02p_pytest_pipe.log 2>&1; tail -15 gpurun_out/r02p_pytest_pipe.log
if ! grep -q "passed" gpurun_out/r02p_pytest_pipe.log || grep -q "failed" gpurun_out/r02p_pytest_pipe.log; then echo "PARITY NOT GREEN: no timing"; exit 1; fi
for v in "NW_PIPE=0" "NW_PIPE=32 NW_TILE_NODES=136" "NW_PIPE=32 NW_TILE_NODES=128" "NW_PIPE=32 NW_TILE_NODES=112" "NW_PIPE=43 NW_TILE_NODES=136" "NW_PIPE=43 NW_TILE_NODES=128" "NW_PIPE=22 NW_TILE_NODES=136" "NW_PIPE=22 NW_TILE_NODES=128"; do
  echo "=== [$v]"
  env $v timeout 150 python bench.py --steps 20 --warmup 5 --detail --sst --no-cpu-baseline 2>&1 >/dev/null | grep -E "ms x" | grep -E "momentum_uvw|continuity|mdot|grad_vector|scalar   " | awk '{printf "%s %s  ", $1, $2} END {print ""}'
done
