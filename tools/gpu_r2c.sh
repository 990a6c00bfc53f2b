#!/bin/bash
# Round 2 kernel iteration (1 GPU): GPU tests, bench --detail for a list of
# environment variants, phase cycles of the default.
#   gpurun --timeout 1200 -- 'bash tools/gpu_r2c.sh TAG "VAR=1 VAR2=2" "VAR=3" ...'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${1:-r02c}; shift || true
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_gpu.log
i=0
for v in "$@"; do
  name=${TAG}_bench_v$i
  echo "=== bench variant $i: [$v]"
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --detail --sst --no-cpu-baseline > gpurun_out/$name.json 2> gpurun_out/$name.detail.txt
  echo "# variant: [$v]" >> gpurun_out/$name.detail.txt
  grep "ms x" gpurun_out/$name.detail.txt
  python -c "
import json;d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'sweep_frac',round(d['roofline']['sweep_frac'],4),'e2e',round(d['e2e']['value'],1))"
  i=$((i+1))
done
echo "=== phase cycles"
make -C nalu-wind_b200 -s prof > /dev/null 2>&1 && NW_LIB_PATH=$PWD/nalu-wind_b200/libnalu_edge_b200_prof.so timeout 300 python tools/phase_times.py > gpurun_out/${TAG}_phase_cycles.txt 2>&1
grep -A8 -E "^(continuity|momentum|mdot|grad_scalar) " gpurun_out/${TAG}_phase_cycles.txt
