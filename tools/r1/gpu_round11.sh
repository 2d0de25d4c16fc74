#!/bin/bash
# Round-1 session 4, second pass: parity of the Gaussian kernels (rows: radius 1-8, scalar and packed column taps), the
# rows / rows_packed / stream sweep over radius 1-8, one ncu capture of the packed rows kernel (radius 4 and 8).
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "== parity: Gaussian kernels =="
timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gaussian" 2>&1 | tail -12 | tee gpurun_out/r11_parity_gauss.log
el "== sweep: rows / rows_packed / stream, radius 1-8 =="
rm -f gpurun_out/kernel_sweep.jsonl
timeout 120 python tools/bench_kernels.py gauss_rows 2>&1 | tail -2
cp gpurun_out/kernel_sweep.jsonl gpurun_out/r11_gauss_rows_sweep.jsonl 2>/dev/null
el "== ncu --set full: k_gauss_rows (packed) radius 4 and 8 at 8192^2 =="
timeout 60 ncu --set full --clock-control none --import-source on -k regex:'k_gauss_rows' -s 2 -c 2 -f -o gpurun_out/prof_gauss_rows_packed \
    python tools/bench_kernels.py gauss_rows_ncu > gpurun_out/ncu_gauss_rows_packed.log 2>&1
el "done"
