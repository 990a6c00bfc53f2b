#!/bin/bash
# multi-GPU round: parity (peer-memory path and NCCL path) + bench at N ranks (N = $1, default 2)
set -u
N=${1:-2}
QUICK=${2:-0}   # 1: peer-memory legs only
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MODES="1 0"; [ "$QUICK" = 1 ] && MODES="1"
for p2p in $MODES; do
for per in 0 1; do
echo "=== mgpu parity periodic=$per p2p=$p2p"
NW_P2P=$p2p NW_MGPU_PERIODIC=$per timeout 300 $TR --master-port 295$p2p$per tests/mgpu_parity.py > gpurun_out/mgpu_parity_p${per}_p2p$p2p.log 2>&1; tail -1 gpurun_out/mgpu_parity_p${per}_p2p$p2p.log | cut -c1-300
done
done
for p2p in $MODES; do
echo "=== bench N=$N p2p=$p2p"
NW_P2P=$p2p timeout 600 $TR --master-port 2952$p2p bench.py --gpus $N --steps 10 --warmup 3 --detail > gpurun_out/bench_n${N}_p2p$p2p.json 2> gpurun_out/bench_n${N}_p2p$p2p.err; grep "ms x" gpurun_out/bench_n${N}_p2p$p2p.err; cut -c1-230 gpurun_out/bench_n${N}_p2p$p2p.json
done
[ "$QUICK" = 1 ] && exit 0
echo "=== bench N=$N sst"
timeout 600 $TR --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 --detail --sst > gpurun_out/bench_n${N}_sst.json 2> gpurun_out/bench_n${N}_sst.err; grep "ms x" gpurun_out/bench_n${N}_sst.err; cut -c1-230 gpurun_out/bench_n${N}_sst.json
