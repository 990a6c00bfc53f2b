#!/bin/bash
# N-GPU bench line only (parity gate + weak headline + north-star record)
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_r2x.sh 8 TAG'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-8}
TAG=${2:-r02x}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 --detail \
  > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "rc=$?"
grep -E "ms x|Error|error" gpurun_out/${TAG}_bench_n$N.err | tail -30
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "sustained", d["sustained"]["value"], "exchange", d.get("halo_exchange"))
print("gate", d.get("parity_gate"))
ns = d.get("north_star", {})
print("north_star", {k: ns.get(k) for k in ("ms_per_sweep", "ms_per_sweep_without_exchanges", "gedges_per_s_per_gpu", "sweep_frac", "exchange_share", "error")})
PY
