#!/bin/bash
# launch list of rank 0 inside the N-rank sweep (kernel durations beside N=1's
# profiles/r02n_launches.csv); a number printed under ncu is not a bench value
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_r2u.sh 2'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
echo "=== ncu launch list of rank 0, N=$N"
NCU_LOG=gpurun_out/r02u_launches_n${N}_rank0.csv NW_BENCH_NORTH_STAR=off NW_P2P_TIMEOUT_S=120 timeout 700 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29871 bash tools/ncu_rank0.sh bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --sustain-s 0.001 > gpurun_out/r02u_ncu_n${N}.log 2>&1
echo rc=$?; tail -2 gpurun_out/r02u_ncu_n${N}.log | cut -c1-300
python - $N <<'PY'
import csv,sys
n=sys.argv[1]
rows=[r for r in csv.reader(l for l in open("gpurun_out/r02u_launches_n%s_rank0.csv" % n) if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
seq=[(r[ki][:70], float(r[vi].replace(",","")), r[ui]) for r in rows[1:]]
print(len(seq),"kernels; last 70:")
for k,v,u in seq[-70:]:
    print("  %-70s %10.1f %s"%(k,v,u))
PY
