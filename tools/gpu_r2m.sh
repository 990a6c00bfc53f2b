#!/bin/bash
# 2-GPU check: partitioned parity (plain + periodic box, peer memory and NCCL) + a short bench line
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29700
for p2p in 1 0; do
for per in 0 1; do
port=$((port+1))
echo "=== mgpu parity periodic=$per p2p=$p2p"
NW_P2P=$p2p NW_MGPU_PERIODIC=$per timeout 300 $TR --master-port $port tests/mgpu_parity.py > gpurun_out/r02m_mgpu2_parity_p${per}_p2p$p2p.json 2> gpurun_out/r02m_mgpu2_parity_p${per}_p2p$p2p.err
echo rc=$?; tail -1 gpurun_out/r02m_mgpu2_parity_p${per}_p2p$p2p.json | cut -c1-330
done
done
port=$((port+1))
echo "=== bench N=2 (north star off)"
NW_BENCH_NORTH_STAR=off timeout 400 $TR --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 --detail > gpurun_out/r02m_bench_n2.json 2> gpurun_out/r02m_bench_n2.detail.txt
grep "ms x" gpurun_out/r02m_bench_n2.detail.txt; cut -c1-200 gpurun_out/r02m_bench_n2.json
