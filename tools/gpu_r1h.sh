#!/bin/bash
# round-1h GPU call: parity incl. 2-D quad O-grid, bench default / SST, 384-thread CTA experiment
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
echo "=== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
run() { name=$1; shift
  echo "=== bench $name: $*"
  timeout 400 python bench.py --detail "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  grep "ms x" gpurun_out/bench_$name.err; python - <<P
import json
d=json.load(open("gpurun_out/bench_$name.json"))
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "sweep_frac", d["roofline"]["sweep_frac"], d.get("cpu_baseline",{}).get("value"))
P
}
run default
run sst --sst --no-cpu-baseline
NW_TILE_THREADS=384 run sst_thr384 --sst --no-cpu-baseline
NW_TILE_THREADS=384 run thr384_t256 --tile 256 --no-cpu-baseline
run t256 --tile 256 --no-cpu-baseline
