#!/bin/bash
# FFMA2 issue-rate microbenchmark + single-GPU parity + bench of the current tree
mkdir -p gpurun_out
echo "== ffma2 microbench =="
timeout 120 tools/microbench/ffma2 2>&1 | tee gpurun_out/ffma2_microbench.txt
echo "== parity =="
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== bench =="
timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r6.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels'], d['diffusion']['gbs'])"
