#!/bin/bash
# A/B of staging variants inside one call (same box, same clocks), 3 repeats each
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-ab}; shift || true
for rep in 1 2; do
for v in "$@"; do
  echo "=== rep $rep [$v]"
  env $v timeout 300 python bench.py --steps 30 --warmup 5 --detail --sst --no-fuse-scalars --no-cpu-baseline 2>&1 >/dev/null | grep -E "ms x" | grep -E "momentum_uvw|continuity|mdot|scalar   " | awk '{printf "%s %s  ", $1, $2} END {print ""}'
done
done
