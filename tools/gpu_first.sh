#!/bin/bash
# first GPU contact: build check, sanitizer on a tiny case, parity tests, bench, ncu
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "=== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'oracle')
import __graft_entry__ as g, parity_util as pu
P=g.load_package(); ctx=P.Context(0)
print(pu.run_lowmach_case(P, ctx, dims=(6,5,4), tile_nodes=32))
print(pu.run_lowmach_case(P, ctx, dims=(6,5,4), tile_nodes=32, mode=1))
" 2>&1 | tail -25 | tee gpurun_out/sanitizer.log
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "=== bench"; for t in 0 128 384; do
  timeout 600 python bench.py --steps 10 --warmup 3 --tile $t --detail --no-cpu-baseline > gpurun_out/bench_tile$t.json 2> gpurun_out/bench_tile$t.err
  tail -12 gpurun_out/bench_tile$t.err; cut -c1-600 gpurun_out/bench_tile$t.json
done
echo "=== bench atomic"; timeout 600 python bench.py --steps 10 --warmup 3 --mode atomic --detail --no-cpu-baseline > gpurun_out/bench_atomic.json 2> gpurun_out/bench_atomic.err; tail -8 gpurun_out/bench_atomic.err
echo "=== bench default (with cpu baseline)"; timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -3 gpurun_out/ncu_launches.log
echo "=== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ls_tile_kernel -s 6 -c 2 -o gpurun_out/prof_ls_tile python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
