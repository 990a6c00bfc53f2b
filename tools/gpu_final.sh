#!/bin/bash
# end-of-round GPU call: smoke, full parity suite, default bench (+ SST), launch list, memcheck of the smoke
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
echo "=== smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log | cut -c1-160
echo "=== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
echo "=== bench default"; timeout 400 python bench.py --detail > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; grep "ms x" gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
echo "=== bench sst"; timeout 400 python bench.py --detail --sst --no-cpu-baseline > gpurun_out/bench_sst.json 2> gpurun_out/bench_sst.err; grep "ms x" gpurun_out/bench_sst.err; cut -c1-200 gpurun_out/bench_sst.json
echo "=== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-100
echo "=== memcheck smoke"; timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1; tail -4 gpurun_out/memcheck_smoke.log | cut -c1-200
