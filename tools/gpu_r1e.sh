#!/bin/bash
# round-1e GPU call: parity (incl. tet-split, staged upload), default bench (fused Peclet, pipelined e2e),
# comparison legs, SST sweep, reference arm, ncu launch list + full capture
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
echo "=== smoke"; timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log | cut -c1-200
echo "=== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
run() { name=$1; shift
  echo "=== bench $name: $*"
  timeout 400 python bench.py --detail "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  tail -7 gpurun_out/bench_$name.err; python - <<P
import json
d=json.load(open("gpurun_out/bench_$name.json"))
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "sweep_frac", d["roofline"]["sweep_frac"], d.get("cpu_baseline",{}).get("value"))
P
}
run default
run serial_upload --serial-upload --no-cpu-baseline
run nofuse --no-fuse-peclet --no-cpu-baseline
run sst --sst --no-cpu-baseline
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-200 gpurun_out/bench_reference.json
echo "=== ncu launch list"; timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-120
echo "=== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_kernel" -s 15 -c 5 -o gpurun_out/prof_r1e python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out | head -40
