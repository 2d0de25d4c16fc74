#!/bin/bash
echo "== parity (stats kernel etc.) =="
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for lib in libslime_b200.so libslime_b200_mb8.so; do
  echo "== $lib =="
  SM_LIB_PATH=$PWD/slime_mold_b200/$lib timeout 600 python bench.py --steps 600 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), d['ms_per_step'], d['kernels'])"
done
