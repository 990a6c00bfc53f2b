#!/bin/bash
# round-1d GPU call: tile kernel with overlapped TMA issue / halo gather, fused peclet, 3-CTA builds, stream variant
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
echo "=== pytest gpu"; timeout 600 python -m pytest tests -x -q -m gpu --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
echo "=== pytest gpu (stream variant)"; NW_STREAM=1 timeout 600 python -m pytest tests -x -q -m gpu --timeout 240 > gpurun_out/pytest_gpu_stream.log 2>&1; tail -3 gpurun_out/pytest_gpu_stream.log
run() { # name, env..., -- args
  name=$1; shift
  echo "=== bench $name"
  env "$@" timeout 300 python bench.py --detail --no-cpu-baseline $ARGS > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  tail -7 gpurun_out/bench_$name.err; cut -c1-140 gpurun_out/bench_$name.json
}
ARGS="--tile 192" run tile192 NW_STREAM=0
ARGS="--tile 192 --fuse-peclet" run tile192_fused NW_STREAM=0
ARGS="--tile 160" run tile160 NW_STREAM=0
ARGS="--tile 112" run tile112_c3 NW_TILE_CTAS=3
ARGS="--tile 96" run tile96_c3 NW_TILE_CTAS=3
ARGS="--tile 112" run stream112_c3 NW_STREAM=1 NW_STREAM_CTAS=3
echo "=== phases"; timeout 200 python tools/phase_times.py > gpurun_out/phases_tile192.txt 2>&1; cat gpurun_out/phases_tile192.txt | head -30
