#!/bin/bash
# Round-1 session 4: parity of every Gaussian kernel (rows / stream / tile / two-pass), the engine against the vectors
# produced from the reference's shader source (tests/test_gpu_wgsl.py), the rest of the GPU suite, the rows-vs-stream
# sweep, the headline bench line, one ncu capture of the register-streaming kernel.
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "== parity: engine vs the reference's shader source =="
timeout 60 python -m pytest tests/test_gpu_wgsl.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r10_parity_wgsl.log
el "== parity: Gaussian kernels =="
timeout 90 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gaussian" 2>&1 | tail -12 | tee gpurun_out/r10_parity_gauss.log
el "== sweep: rows vs stream vs tile =="
rm -f gpurun_out/kernel_sweep.jsonl
timeout 90 python tools/bench_kernels.py gauss_rows 2>&1 | tail -2
cp gpurun_out/kernel_sweep.jsonl gpurun_out/r10_gauss_rows_sweep.jsonl 2>/dev/null
el "== the rest of the GPU suite =="
timeout 120 python -m pytest tests -q -m gpu -k "not gaussian and not wgsl" 2>&1 | tail -6 | tee gpurun_out/r10_parity_rest.log
el "== bench.py (headline) =="
timeout 120 python bench.py 2>&1 | tail -1 > gpurun_out/r10_bench.log
python -c "import json; d=json.load(open('gpurun_out/r10_bench.log')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
el "== ncu --set full: k_gauss_rows radius 2 and 4 at 8192^2 =="
timeout 60 ncu --set full --clock-control none --import-source on -k regex:'k_gauss_rows' -s 2 -c 2 -f -o gpurun_out/prof_gauss_rows \
    python tools/bench_kernels.py gauss_rows_ncu > gpurun_out/ncu_gauss_rows.log 2>&1
el "done"
