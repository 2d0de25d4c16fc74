#!/bin/bash
# parity + bench (Default / Snake / Waves) of the current tree
mkdir -p gpurun_out
echo "== parity =="
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for p in Default Snake Waves; do
echo "== bench $p =="
timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --preset $p 2>&1 | tail -1 | tee gpurun_out/bench_r7_$p.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels'], d['diffusion']['gbs'])"
done
