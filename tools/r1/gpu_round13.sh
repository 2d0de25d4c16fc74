#!/bin/bash
# Round-1 closing pass (90 s of GPU budget): parity of every Gaussian variant (incl. the FFMA2 forms) and of the engine
# against the shader-source vectors, the packed A/B sweep, the rest of the GPU suite, smoke().
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "== parity: shader-source vectors + Gaussian kernels =="
timeout 60 python -m pytest tests/test_gpu_wgsl.py tests/test_gpu_parity.py -q -m gpu -k "wgsl or gaussian" 2>&1 | tail -8 | tee gpurun_out/r13_parity_gauss_wgsl.log
el "== sweep: FFMA2 forms =="
rm -f gpurun_out/kernel_sweep.jsonl
timeout 40 python tools/bench_kernels.py gauss_packed 2>&1 | tail -1
cp gpurun_out/kernel_sweep.jsonl gpurun_out/r13_gauss_packed_sweep.jsonl 2>/dev/null
el "== the rest of the GPU suite =="
timeout 60 python -m pytest tests -q -m gpu -k "not gaussian and not wgsl" 2>&1 | tail -5 | tee gpurun_out/r13_parity_rest.log
el "== smoke =="
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r13_smoke.log
el "done"
