#!/bin/bash
# N-GPU check of the fused halo exchange: partitioned parity (incl. the
# eager-exchange bit-identity checks of tests/mgpu_parity.py) on both transports
# and with the asynchronous pulls, then the weak-scaling bench line
#   fused + asynchronous pulls | fused | plain (NW_HALO_OVERLAP=0)
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r2t.sh [N]'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29800
for cfg in "1 0 0" "1 1 1" "1 0 1" "0 0 0"; do
set -- $cfg; p2p=$1; per=$2; as=$3
port=$((port+1))
name=r02t_mgpu${N}_parity_p${per}_p2p${p2p}_async${as}
echo "=== mgpu parity periodic=$per p2p=$p2p async=$as"
NW_P2P=$p2p NW_P2P_ASYNC=$as NW_MGPU_PERIODIC=$per timeout 300 $TR --master-port $port tests/mgpu_parity.py > gpurun_out/$name.json 2> gpurun_out/$name.err
echo rc=$?; tail -1 gpurun_out/$name.json | cut -c1-420
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/$name.err | tail -3
done
for cfg in "1 1" "1 0" "0 0"; do
set -- $cfg; ov=$1; as=$2
for sst in "" "--sst"; do
port=$((port+1))
name=r02t_bench_n${N}_overlap${ov}_async${as}${sst:+_sst}
echo "=== bench N=$N overlap=$ov async=$as $sst (north star off)"
NW_HALO_OVERLAP=$ov NW_P2P_ASYNC=$as NW_BENCH_NORTH_STAR=off timeout 400 $TR --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --detail --no-cpu-baseline $sst > gpurun_out/$name.json 2> gpurun_out/$name.detail.txt
grep "ms x" gpurun_out/$name.detail.txt
python -c "
import json;d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'sustained',round(d['sustained']['value'],1),'parity',d.get('parity_gate',{}).get('ok'),'exchange',d.get('halo_exchange'))"
done
done
