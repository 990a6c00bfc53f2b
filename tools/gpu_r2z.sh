#!/bin/bash
# monolithic momentum on the tile path: its tests + the --detail lines
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q -k "monolithic or extract_diagonal or real_mesh or fuzz" 2>&1 | tail -5
timeout 400 python bench.py --steps 10 --warmup 3 --detail --no-cpu-baseline > gpurun_out/r02z_bench_default.json 2> gpurun_out/r02z_bench_default.detail.txt
grep "ms x\|mono" gpurun_out/r02z_bench_default.detail.txt
