#!/bin/bash
# what the driver runs at round end for N GPUs: bench.py with default flags
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29931 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02final_bench_n$N.json 2> gpurun_out/r02final_bench_n$N.err
echo "rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02final_bench_n$N.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","unit","n_gpus","ms_per_step","scaling","gpu_launches")})
print("e2e", d["e2e"]["value"], "sustained", d["sustained"]["value"], "exchange", d.get("halo_exchange"))
print("gate", d.get("parity_gate", {}).get("ok"), "config.exchange:", d["config"].get("exchange"))
ns = d.get("north_star", {})
print("north_star", {k: ns.get(k) for k in ("ms_per_sweep", "gedges_per_s_per_gpu", "sweep_frac", "exchange_share", "error", "skipped")})
PY
tail -3 gpurun_out/r02final_bench_n$N.err | cut -c1-300
