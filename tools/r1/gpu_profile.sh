#!/bin/bash
# Profiling pass for profiles/: launch list, full ncu captures (cold + in situ), kernel sweeps, bench lines.
mkdir -p gpurun_out
rm -f gpurun_out/kernel_sweep.jsonl
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1; nproc > gpurun_out/nproc.txt
echo "== bench (default) =="
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.log
echo "== bench --impl reference =="
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_reference.log
echo "== launch list =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 40 --warmup 3 --spinup 40 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
echo "== ncu full (cold cache) =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 50 -c 2 -f -o gpurun_out/prof_agents \
    python bench.py --steps 10 --warmup 2 --spinup 40 --no-cpu-baseline > gpurun_out/ncu_agents.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_trail_rows' -s 50 -c 2 -f -o gpurun_out/prof_trail \
    python bench.py --steps 10 --warmup 2 --spinup 40 --no-cpu-baseline > gpurun_out/ncu_trail.log 2>&1
echo "== ncu full (in situ, cache-control none) =="
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_agents|k_trail_rows|k_tile' -s 60 -c 10 -f -o gpurun_out/prof_insitu \
    python bench.py --steps 24 --warmup 2 --spinup 30 --no-cpu-baseline > gpurun_out/ncu_insitu.log 2>&1
echo "== ncu diffusion 16384^2 =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_trail_rows' -s 8 -c 2 -f -o gpurun_out/prof_diffusion \
    python tools/bench_kernels.py diffusion16k > gpurun_out/ncu_diffusion.log 2>&1
echo "== sweeps =="
timeout 1500 python tools/bench_kernels.py diffusion agents presets 2>&1 | tail -5
echo "== config 1 and config 3 =="
timeout 600 python bench.py --agents 1000000 --width 1920 --height 1080 --steps 1000 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_config1.log
timeout 900 python bench.py --agents 100000000 --width 8192 --height 8192 --preset Snake --steps 100 --warmup 5 --spinup 100 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_config3.log
ls -la gpurun_out
