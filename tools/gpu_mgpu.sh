#!/bin/bash
# multi-GPU round: parity + bench at N ranks for the halo transports
#   $1 = N (default 2); $2 = 1: default transport only (peer memory, asynchronous pull)
set -u
N=${1:-2}
QUICK=${2:-0}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# mode = <NW_P2P><NW_P2P_ASYNC>: 11 peer memory + async pull (default), 10 peer memory one stream, 00 NCCL
MODES="11 10 00"; [ "$QUICK" = 1 ] && MODES="11"
port=29500
for m in $MODES; do
p2p=${m:0:1}; as=${m:1:1}
for per in 0 1; do
port=$((port+1))
echo "=== mgpu parity periodic=$per p2p=$p2p async=$as"
NW_P2P=$p2p NW_P2P_ASYNC=$as NW_MGPU_PERIODIC=$per timeout 300 $TR --master-port $port tests/mgpu_parity.py > gpurun_out/mgpu_parity_p${per}_m$m.log 2>&1; tail -1 gpurun_out/mgpu_parity_p${per}_m$m.log | cut -c1-260
done
done
for m in $MODES; do
p2p=${m:0:1}; as=${m:1:1}
port=$((port+1))
echo "=== bench N=$N p2p=$p2p async=$as"
NW_P2P=$p2p NW_P2P_ASYNC=$as timeout 600 $TR --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --detail > gpurun_out/bench_n${N}_m$m.json 2> gpurun_out/bench_n${N}_m$m.err; grep "ms x" gpurun_out/bench_n${N}_m$m.err; cut -c1-230 gpurun_out/bench_n${N}_m$m.json
done
[ "$QUICK" = 1 ] && exit 0
port=$((port+1))
echo "=== bench N=$N sst"
timeout 600 $TR --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --detail --sst > gpurun_out/bench_n${N}_sst.json 2> gpurun_out/bench_n${N}_sst.err; grep "ms x" gpurun_out/bench_n${N}_sst.err; cut -c1-230 gpurun_out/bench_n${N}_sst.json
