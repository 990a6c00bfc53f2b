#!/bin/bash
# GPU round: smoke, parity tests, bench variants, ncu launch list + full capture
# usage: gpu_round.sh "<tile sizes>" [skip-tests]
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
TILES=${1:-"0"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
echo "=== smoke"; timeout 120 python -X faulthandler -c "import faulthandler; faulthandler.dump_traceback_later(100, exit=True); import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log | cut -c1-400
if [ "${2:-}" != "skip-tests" ]; then
echo "=== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
fi
echo "=== bench"; for t in $TILES; do
  timeout 300 python bench.py --steps 10 --warmup 3 --tile $t --detail --no-cpu-baseline > gpurun_out/bench_tile$t.json 2> gpurun_out/bench_tile$t.err
  echo "tile $t"; tail -12 gpurun_out/bench_tile$t.err; cut -c1-120 gpurun_out/bench_tile$t.json
done
echo "=== bench atomic"; timeout 300 python bench.py --steps 10 --warmup 3 --mode atomic --detail --no-cpu-baseline > gpurun_out/bench_atomic.json 2> gpurun_out/bench_atomic.err; tail -8 gpurun_out/bench_atomic.err
echo "=== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --tile ${NCU_TILE:-0} --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log | cut -c1-200
echo "=== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 24 -c 6 -o gpurun_out/prof_tiles python bench.py --steps 2 --warmup 3 --tile ${NCU_TILE:-0} --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out
