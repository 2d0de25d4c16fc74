#!/bin/bash
# Round 2, GPU call 2: microbenchmarks that decide the layouts (incoherent footprint rate per layout; surface read/write rates).
mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el gather_rate 2048; timeout 120 tools/microbench/gather_rate 2048 2>&1 | tee gpurun_out/r2_micro_gather_2048.log
el gather_rate 1024; timeout 120 tools/microbench/gather_rate 1024 2>&1 | tee gpurun_out/r2_micro_gather_1024.log
el surfcopy; timeout 120 tools/microbench/surfcopy 2>&1 | tee gpurun_out/r2_micro_surfcopy.log
el "parity after the ADVICE fixes (subset)"
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wgsl.py -q -m gpu -x 2>&1 | tail -4
el done
