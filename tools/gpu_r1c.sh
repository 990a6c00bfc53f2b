#!/bin/bash
# round-1c GPU call: parity state + bench + per-phase cycle accounting + ncu full (stall reasons)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
echo "=== smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log | cut -c1-300
echo "=== pytest gpu"; timeout 600 python -m pytest tests -x -q -m gpu --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
echo "=== bench"; timeout 400 python bench.py --detail > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -8 gpurun_out/bench_default.err; cut -c1-200 gpurun_out/bench_default.json
echo "=== phases"; timeout 200 python tools/phase_times.py > gpurun_out/phases.txt 2>&1; cat gpurun_out/phases.txt
echo "=== ncu full"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:"ls_tile_kernel|mdot_tile" -s 7 -c 3 -o gpurun_out/prof_r1c python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out
