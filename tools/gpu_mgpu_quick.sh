#!/bin/bash
# quick 2-rank A/B of the asynchronous pull (bench only)
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
for m in 11 10; do
p2p=${m:0:1}; as=${m:1:1}
port=$((port+1))
echo "=== bench N=$N p2p=$p2p async=$as"
NW_P2P=$p2p NW_P2P_ASYNC=$as timeout 300 $TR --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --detail > gpurun_out/bench_n${N}_m$m.json 2> gpurun_out/bench_n${N}_m$m.err; grep "ms x" gpurun_out/bench_n${N}_m$m.err; cut -c1-230 gpurun_out/bench_n${N}_m$m.json
done
