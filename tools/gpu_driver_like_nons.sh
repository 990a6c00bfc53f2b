#!/bin/bash
# bench.py with default flags except the north-star record (GPU-minute budget)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
t0=$(date +%s)
NW_BENCH_NORTH_STAR=off timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29941 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02final_bench_n${N}_nons.json 2> gpurun_out/r02final_bench_n${N}_nons.err
echo "rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02final_bench_n${N}_nons.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","unit","n_gpus","ms_per_step","scaling","gpu_launches")})
print("e2e", d["e2e"]["value"], "sustained", d["sustained"]["value"], "exchange", d.get("halo_exchange"), "gate", d.get("parity_gate", {}).get("ok"))
PY
