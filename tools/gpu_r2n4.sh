#!/bin/bash
# N-GPU: partitioned parity (periodic box, peer memory, asynchronous pulls) and
# the weak-scaling bench line without the north-star record
#   gpurun --gpus 4 --timeout 600 -- 'bash tools/gpu_r2n4.sh 4'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NW_MGPU_PERIODIC=1 NW_MGPU_DIMS=14,12,28 timeout 300 $TR --master-port 29911 tests/mgpu_parity.py > gpurun_out/r02n4_mgpu${N}_parity.json 2> gpurun_out/r02n4_mgpu${N}_parity.err
echo rc=$?; tail -1 gpurun_out/r02n4_mgpu${N}_parity.json | cut -c1-420
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02n4_mgpu${N}_parity.err | tail -3
NW_BENCH_NORTH_STAR=off timeout 400 $TR --master-port 29912 bench.py --gpus $N --steps 20 --warmup 5 --detail --no-cpu-baseline > gpurun_out/r02n4_bench_n$N.json 2> gpurun_out/r02n4_bench_n$N.detail.txt
grep "ms x" gpurun_out/r02n4_bench_n$N.detail.txt
python -c "
import json;d=json.loads(open('gpurun_out/r02n4_bench_n$N.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],4),'sustained',round(d['sustained']['value'],1),'parity',d.get('parity_gate',{}).get('ok'),'exchange',d.get('halo_exchange'))"
