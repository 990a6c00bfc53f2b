#!/bin/bash
# Round 2 kernel iteration (1 GPU): GPU tests, then bench lines:
#   default (BASELINE configs[1]), SST fused / un-fused scalar pair, the
#   three-edges-per-thread experiment, configs[3] / [4] meshes; phase cycles.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2c.sh TAG'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r02c}
echo "=== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_gpu.log
run() { # name, env, args
  local name=${TAG}_bench_$1
  echo "=== bench $1: [$2] $3"
  env $2 timeout 400 python bench.py --steps 20 --warmup 5 --detail $3 > gpurun_out/$name.json 2> gpurun_out/$name.detail.txt
  echo "# env [$2] args [$3]" >> gpurun_out/$name.detail.txt
  grep "ms x" gpurun_out/$name.detail.txt
  python -c "
import json;d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'sweep_frac',round(d['roofline']['sweep_frac'],4),'mom_frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value'],1),'sustained',round(d['sustained']['value'],1))"
}
run default "" ""
run sst "" "--sst --no-cpu-baseline"
if [ "${2:-}" = "all" ]; then
run sst_fused "NW_SCALAR_PAIR_FUSED=1" "--sst --fuse-scalars --no-cpu-baseline"
run warped "" "--mesh warped --sst --no-cpu-baseline"
run mixed "" "--mesh mixed --sst --no-cpu-baseline"
run mixed_refine "NW_TILE_REFINE=1" "--mesh mixed --sst --no-cpu-baseline"
fi
echo "=== phase cycles"
make -C nalu-wind_b200 -s prof > /dev/null 2>&1 && NW_LIB_PATH=$PWD/nalu-wind_b200/libnalu_edge_b200_prof.so timeout 300 python tools/phase_times.py > gpurun_out/${TAG}_phase_cycles.txt 2>&1
grep -A8 -E "^(continuity|momentum|mdot|grad_scalar) " gpurun_out/${TAG}_phase_cycles.txt
