#!/bin/bash
# Second GPU pass: parity after the kernel trims, bench, kernel sweeps, in-situ ncu.
mkdir -p gpurun_out
echo "== pytest gpu =="
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench =="
timeout 600 python bench.py --steps 2000 --warmup 10 --spinup 200 2>&1 | tail -1 | tee gpurun_out/bench.log
echo "== sweeps =="
rm -f gpurun_out/kernel_sweep.jsonl
timeout 1500 python tools/bench_kernels.py diffusion agents presets 2>&1 | tail -80 | tee gpurun_out/sweeps.log
echo "== ncu (cache-control none, in situ) =="
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_agents|k_trail_rows|k_tile' -s 60 -c 8 -f -o gpurun_out/prof_insitu \
    python bench.py --steps 20 --warmup 2 --spinup 30 --no-cpu-baseline > gpurun_out/ncu_insitu.log 2>&1
ls -la gpurun_out
