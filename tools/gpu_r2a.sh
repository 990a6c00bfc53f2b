#!/bin/bash
# Round 2, first GPU call (8 GPUs, charged 8x: keep it short):
#   1. bench.py at N = 8: parity gate, weak 128^3/GPU headline, sustained leg,
#      e2e, and the north-star record (512^3 SST sweep over the 8 GPUs);
#   2. 8-rank parity on the reference's own decomposition (hybrid.g.8.0-7)
#      against the serial oracle, peer-memory transport, + the delayed-rank test.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_r2a.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-8}
TAG=${2:-r02a}
{
  nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA node|^CPU\(s\)"
  nvidia-smi --query-gpu=index,name,memory.total --format=csv
  nvidia-smi topo -m
} > gpurun_out/${TAG}_box.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== bench N=$N (north star)"
t0=$(date +%s)
timeout 700 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --detail \
  > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "rc=$? wall=$(( $(date +%s) - t0 ))s"
grep -E "ms x|\[bench\]|Error|error" gpurun_out/${TAG}_bench_n$N.err | tail -30
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "sweep_frac", d["roofline"]["sweep_frac"], "e2e", d["e2e"]["value"])
    print("gate", d.get("parity_gate"))
    ns = d.get("north_star", {})
    print("north_star", {k: ns.get(k) for k in ("ms_per_sweep", "gedges_per_s_per_gpu", "sweep_frac", "exchange_share", "setup_seconds", "skipped", "error")})
except Exception as e:
    print("no bench line:", e)
PY
if [ "$N" = 8 ]; then
echo "=== 8-rank parity on hybrid.g.8 (peer memory) + delayed-rank test"
t0=$(date +%s)
NW_MGPU_MESH=hybrid8 NW_MGPU_DELAY=1 NW_P2P_TIMEOUT_S=2 timeout 300 $TR --master-port 29542 tests/mgpu_parity.py \
  > gpurun_out/${TAG}_mgpu8_parity.json 2> gpurun_out/${TAG}_mgpu8_parity.err
echo "rc=$? wall=$(( $(date +%s) - t0 ))s"
tail -1 gpurun_out/${TAG}_mgpu8_parity.json | cut -c1-900
tail -5 gpurun_out/${TAG}_mgpu8_parity.err
fi
