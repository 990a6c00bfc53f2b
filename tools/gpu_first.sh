#!/bin/bash
# GPU round: smoke, parity tests, bench variants, ncu launch list + full capture
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
echo "=== smoke"; timeout 120 python -X faulthandler -c "import faulthandler; faulthandler.dump_traceback_later(100, exit=True); import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
echo "=== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
echo "=== bench"; for t in 0 128 512; do
  timeout 300 python bench.py --steps 10 --warmup 3 --tile $t --detail --no-cpu-baseline > gpurun_out/bench_tile$t.json 2> gpurun_out/bench_tile$t.err
  tail -12 gpurun_out/bench_tile$t.err; cut -c1-300 gpurun_out/bench_tile$t.json
done
echo "=== bench atomic"; timeout 300 python bench.py --steps 10 --warmup 3 --mode atomic --detail --no-cpu-baseline > gpurun_out/bench_atomic.json 2> gpurun_out/bench_atomic.err; tail -8 gpurun_out/bench_atomic.err
echo "=== bench default (with cpu baseline)"; timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
echo "=== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
echo "=== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -3 gpurun_out/ncu_launches.log
echo "=== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:ls_tile_kernel -s 6 -c 2 -o gpurun_out/prof_ls_tile python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
