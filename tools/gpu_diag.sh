#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 120 python tools/diag_realmesh.py > gpurun_out/diag_realmesh.log 2>&1; tail -30 gpurun_out/diag_realmesh.log
