#!/bin/bash
# first GPU contact: build check, sanitizer on a tiny case, parity tests, bench
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
( time timeout 600 python -c "import __graft_entry__ as g; g.build(); print('build ok')" ) 2>&1 | tail -5 | tee gpurun_out/build.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "=== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'oracle')
import __graft_entry__ as g, parity_util as pu
P=g.load_package(); ctx=P.Context(0)
print(pu.run_lowmach_case(P, ctx, dims=(6,5,4), tile_nodes=32))
print(pu.run_lowmach_case(P, ctx, dims=(6,5,4), tile_nodes=32, mode=1))
" 2>&1 | tail -25 | tee gpurun_out/sanitizer.log
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "=== bench"; for t in 0 128 192 384; do
  timeout 600 python bench.py --steps 10 --warmup 3 --tile $t --detail --no-cpu-baseline > gpurun_out/bench_tile$t.json 2> gpurun_out/bench_tile$t.err
  tail -12 gpurun_out/bench_tile$t.err; cut -c1-400 gpurun_out/bench_tile$t.json
done
echo "=== bench atomic"; timeout 600 python bench.py --steps 10 --warmup 3 --mode atomic --detail --no-cpu-baseline > gpurun_out/bench_atomic.json 2> gpurun_out/bench_atomic.err; tail -8 gpurun_out/bench_atomic.err
