#!/bin/bash
# N = 2 with the north-star record (512^3 over two GPUs): memory / time check of
# what the driver's scaling run does at N = 2
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
t0=$(date +%s)
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29761 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02q_bench_n$N.json 2> gpurun_out/r02q_bench_n$N.err
echo "rc=$? wall=$(( $(date +%s) - t0 ))s"
tail -3 gpurun_out/r02q_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02q_bench_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "gate", d.get("parity_gate", {}).get("ok"))
ns = d.get("north_star", {})
print({k: ns.get(k) for k in ("ms_per_sweep", "gedges_per_s_per_gpu", "sweep_frac", "exchange_share", "setup_seconds", "skipped", "error")})
PY
nvidia-smi --query-gpu=memory.used --format=csv | head -3
free -g | head -2
