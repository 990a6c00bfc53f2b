#!/bin/bash
# quick multi-rank parity (peer-memory and NCCL transports) -- no bench
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
for m in ${MODES:-1 0}; do
for per in 0 1; do
port=$((port+1))
echo "=== mgpu parity periodic=$per p2p=$m"
NW_P2P=$m NW_MGPU_PERIODIC=$per timeout 300 $TR --master-port $port tests/mgpu_parity.py > gpurun_out/mgpu_parity_p${per}_p2p$m.log 2>&1; tail -1 gpurun_out/mgpu_parity_p${per}_p2p$m.log | cut -c1-900
done
done
