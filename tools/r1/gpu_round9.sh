#!/bin/bash
# Short follow-up pass: parity of the Gaussian kernels, stream-vs-tile sweep, one ncu capture of the streaming kernel.
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "== parity: Gaussian kernels =="
timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gaussian" 2>&1 | tail -8 | tee gpurun_out/r9_parity_gauss.log
el "== sweep: stream vs tile =="
rm -f gpurun_out/kernel_sweep.jsonl
timeout 150 python tools/bench_kernels.py gauss_stream 2>&1 | tail -2
cp gpurun_out/kernel_sweep.jsonl gpurun_out/r9_gauss_stream_sweep.jsonl 2>/dev/null
el "== ncu --set full: k_gauss_stream radius 8 and 2 at 8192^2 =="
timeout 120 ncu --set full --clock-control none --import-source on -k regex:'k_gauss_stream' -s 2 -c 2 -f -o gpurun_out/prof_gauss_stream_v2 \
    python tools/bench_kernels.py gauss_ncu > gpurun_out/ncu_gauss_stream_v2.log 2>&1
el "done"
