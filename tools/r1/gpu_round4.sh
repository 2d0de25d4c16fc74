#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu =="
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== pytest gpu parity (no flags, ldg) =="
SM_NO_DEPOSIT_FLAGS=1 SM_SAMPLER=ldg timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for f in 0 1; do
  echo "== bench no_flags=$f =="
  SM_NO_DEPOSIT_FLAGS=$f timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels'])"
done
