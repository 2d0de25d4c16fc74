#!/bin/bash
# Round-1 closing pass (short GPU budget): parity of the streaming Gaussian kernel, its sweep against the tile kernel,
# one ncu capture of it, the evict-first A/B on the agent stream, config 3 with the Gaussian extension, then the rest of
# the GPU suite for as long as the box lasts.  Every step writes its result under gpurun_out/ as soon as it has one.
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "== parity: Gaussian kernels (tile / packed / two-pass / stream) =="
timeout 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gaussian" 2>&1 | tail -15 | tee gpurun_out/r8_parity_gauss.log
el "== sweep: stream vs tile =="
rm -f gpurun_out/kernel_sweep.jsonl
timeout 200 python tools/bench_kernels.py gauss_stream 2>&1 | tail -3
cp gpurun_out/kernel_sweep.jsonl gpurun_out/r8_gauss_stream_sweep.jsonl 2>/dev/null
el "== A/B: evict-first hints on the agent stream (config 2, Default) =="
for h in 0 1; do
  SM_AGENT_STREAM_HINT=$h timeout 120 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r8_bench_hint$h.log
  python -c "import json; d=json.load(open('gpurun_out/r8_bench_hint$h.log')); print('hint=$h', d['value'], d['ms_per_step'], d['kernels']['agents']['ms'], d['kernels']['trail']['ms'], d['e2e']['value'])"
done
el "== ncu --set full: k_gauss_stream radius 8 and 2 at 8192^2 =="
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_gauss_stream' -s 2 -c 2 -f -o gpurun_out/prof_gauss_stream \
    python tools/bench_kernels.py gauss_ncu > gpurun_out/ncu_gauss_stream.log 2>&1
el "== config 3 with the Gaussian extension (100 M agents, 8192^2, Snake sensors, radius 8) =="
SM_GAUSS_KERNEL=stream timeout 150 python bench.py --agents 100000000 --width 8192 --height 8192 --preset Snake --gaussian 8 --steps 50 --warmup 5 --spinup 60 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r8_bench_config3_gauss_stream.log | cut -c1-200
SM_GAUSS_KERNEL=tile timeout 150 python bench.py --agents 100000000 --width 8192 --height 8192 --preset Snake --gaussian 8 --steps 50 --warmup 5 --spinup 60 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r8_bench_config3_gauss_tile.log | cut -c1-200
el "== the rest of the GPU suite =="
timeout 600 python -m pytest tests -x -q -m gpu -k "not gaussian" 2>&1 | tail -4 | tee gpurun_out/r8_parity_rest.log
el "done"
