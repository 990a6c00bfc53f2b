#!/bin/bash
# Round 2, second GPU call (1 GPU): the staging-layout / phase-1 microbenchmarks
# written at the end of round 1, the full GPU test-suite (first run of the
# blocked phase-2 walk, the default-option physics paths and the new halo code),
# a --detail bench line and the per-phase cycle counters.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
bash tools/gpu_microbench.sh > gpurun_out/r02b_microbench.log 2>&1
tail -40 gpurun_out/microbench_stage_layouts.txt
tail -30 gpurun_out/microbench_phase1_momentum.txt
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02b_pytest_gpu.log
echo "=== bench default"
timeout 300 python bench.py --steps 20 --warmup 5 --detail > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.detail.txt
grep "ms x" gpurun_out/r02b_bench_default.detail.txt; cut -c1-300 gpurun_out/r02b_bench_default.json
echo "=== bench sst"
timeout 300 python bench.py --steps 20 --warmup 5 --detail --sst --no-cpu-baseline > gpurun_out/r02b_bench_sst.json 2> gpurun_out/r02b_bench_sst.detail.txt
grep "ms x" gpurun_out/r02b_bench_sst.detail.txt
echo "=== phase cycles"
make -C nalu-wind_b200 -s prof > /dev/null 2>&1 && NW_LIB_PATH=$PWD/nalu-wind_b200/libnalu_edge_b200_prof.so timeout 300 python tools/phase_times.py > gpurun_out/r02b_phase_cycles_tile192.txt 2>&1
cat gpurun_out/r02b_phase_cycles_tile192.txt | head -60
