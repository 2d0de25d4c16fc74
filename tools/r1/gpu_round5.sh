#!/bin/bash
mkdir -p gpurun_out
echo "== parity texlin =="
SM_SAMPLER=texlin timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
echo "== parity default =="
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for s in tex texlin ldg; do
  for p in Default Snake; do
  echo "== bench sampler=$s preset=$p =="
  SM_SAMPLER=$s timeout 600 python bench.py --steps 600 --warmup 10 --no-cpu-baseline --preset $p 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels'], d['diffusion']['gbs'])"
  done
done
