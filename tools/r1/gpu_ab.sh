#!/bin/bash
# A/B of library variants: tools/gpu_ab.sh "<lib1> <lib2> ..." "<preset1> <preset2> ..."
LIBS=${1:-"libslime_b200.so"}
PRESETS=${2:-"Default"}
echo "== parity =="
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for lib in $LIBS; do
  for p in $PRESETS; do
  echo "== $lib $p =="
  SM_LIB_PATH=$PWD/slime_mold_b200/$lib timeout 600 python bench.py --steps 400 --warmup 10 --no-cpu-baseline --preset $p 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), d['ms_per_step'], 'agents %.4f trail %.4f sort %.4f'%(d['kernels']['agents']['ms'], d['kernels']['trail']['ms'], d['kernels']['sort_ms_per_step']))"
  done
done
