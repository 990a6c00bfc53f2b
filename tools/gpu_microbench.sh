#!/bin/bash
# staging-layout and phase-1 microbenchmarks on the GPU box (design aid):
#   gpurun --timeout 300 -- 'bash tools/gpu_microbench.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
  -o /tmp/stage_layouts tools/microbench/stage_layouts.cu -lcuda || exit 1
timeout 120 /tmp/stage_layouts 2.1 20 | tee gpurun_out/microbench_stage_layouts.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude -Inalu-wind_b200/csrc \
  -o /tmp/phase1_momentum tools/microbench/phase1_momentum.cu || exit 1
timeout 120 /tmp/phase1_momentum 200 | tee gpurun_out/microbench_phase1_momentum.txt
