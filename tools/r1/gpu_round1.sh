#!/bin/bash
# First GPU pass: parity tests, smoke, a short bench, launch list, full ncu capture of both hot kernels.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest gpu ==" 
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench =="
timeout 600 python bench.py --steps 100 --warmup 5 --spinup 200 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "== ncu launch list =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 2 --spinup 40 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
echo "== ncu full agents =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_agents -s 45 -c 2 -f -o gpurun_out/prof_agents \
    python bench.py --steps 10 --warmup 2 --spinup 40 --no-cpu-baseline > gpurun_out/ncu_agents.log 2>&1
echo "== ncu full trail =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trail_rows -s 45 -c 2 -f -o gpurun_out/prof_trail \
    python bench.py --steps 10 --warmup 2 --spinup 40 --no-cpu-baseline > gpurun_out/ncu_trail.log 2>&1
ls -la gpurun_out
