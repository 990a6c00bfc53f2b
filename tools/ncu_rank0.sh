#!/bin/bash
# torchrun --no-python helper: rank 0 runs under ncu (launch list only: one pass
# per kernel, no replay), the other ranks run plainly
#   NCU_LOG=gpurun_out/x.csv torchrun --no-python ... bash tools/ncu_rank0.sh bench.py ...
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-3000} --csv --log-file "${NCU_LOG:-gpurun_out/ncu_rank0.csv}" python "$@"
else
  exec python "$@"
fi
