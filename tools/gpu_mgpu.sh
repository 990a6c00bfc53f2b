#!/bin/bash
# multi-GPU round: NCCL parity check + bench at N ranks (N = $1, default 2)
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
for per in 0 1; do
echo "=== mgpu parity periodic=$per"
NW_MGPU_PERIODIC=$per timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$per tests/mgpu_parity.py > gpurun_out/mgpu_parity_p$per.log 2>&1; tail -4 gpurun_out/mgpu_parity_p$per.log | cut -c1-600
done
echo "=== bench N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 --detail > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -12 gpurun_out/bench_n$N.err; cut -c1-400 gpurun_out/bench_n$N.json
echo "=== bench N=$N sst"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --detail --sst > gpurun_out/bench_n${N}_sst.json 2> gpurun_out/bench_n${N}_sst.err; tail -12 gpurun_out/bench_n${N}_sst.err; cut -c1-400 gpurun_out/bench_n${N}_sst.json
