#!/bin/bash
# 2-GPU check of the boundary-tiles-first exchange: partitioned parity (incl. the
# eager-exchange bit-identity checks of tests/mgpu_parity.py) on both transports,
# then the weak-scaling bench line with and without the overlap
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r2t.sh [N]'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29800
for cfg in "1 0" "1 1" "0 0"; do
set -- $cfg; p2p=$1; per=$2
port=$((port+1))
echo "=== mgpu parity periodic=$per p2p=$p2p"
NW_P2P=$p2p NW_MGPU_PERIODIC=$per timeout 300 $TR --master-port $port tests/mgpu_parity.py > gpurun_out/r02t_mgpu${N}_parity_p${per}_p2p$p2p.json 2> gpurun_out/r02t_mgpu${N}_parity_p${per}_p2p$p2p.err
echo rc=$?; tail -1 gpurun_out/r02t_mgpu${N}_parity_p${per}_p2p$p2p.json | cut -c1-700
tail -3 gpurun_out/r02t_mgpu${N}_parity_p${per}_p2p$p2p.err
done
for ov in 1 0; do
for sst in "" "--sst"; do
port=$((port+1))
name=r02t_bench_n${N}_overlap${ov}${sst:+_sst}
echo "=== bench N=$N overlap=$ov $sst (north star off)"
NW_HALO_OVERLAP=$ov NW_BENCH_NORTH_STAR=off timeout 400 $TR --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --detail --no-cpu-baseline $sst > gpurun_out/$name.json 2> gpurun_out/$name.detail.txt
grep "ms x" gpurun_out/$name.detail.txt
python -c "
import json;d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'sustained',round(d['sustained']['value'],1),'parity',d.get('parity_gate',{}).get('ok'))"
done
done
