#!/bin/bash
# TEX-gather sampler: parity + A/B against the LDG sampler.
mkdir -p gpurun_out
echo "== pytest gpu (tex default) =="
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== pytest gpu parity (ldg) =="
SM_SAMPLER=ldg timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_ldg.log
for s in tex ldg; do
  echo "== bench sampler=$s =="
  SM_SAMPLER=$s timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$s.log
done
for p in Snake Curls; do
 for s in tex ldg; do
  echo "== bench preset=$p sampler=$s =="
  SM_SAMPLER=$s timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --preset $p 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['kernels'])"
 done
done
echo "== ncu tex =="
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_agents|k_trail_rows' -s 60 -c 4 -f -o gpurun_out/prof_tex \
    python bench.py --steps 20 --warmup 2 --spinup 30 --no-cpu-baseline > gpurun_out/ncu_tex.log 2>&1
ls gpurun_out
