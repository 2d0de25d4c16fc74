#!/bin/bash
# one full ncu capture of the agent kernel (cold cache), config 2 / preset from $1 (default Default)
P=${1:-Default}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 50 -c 2 -f -o gpurun_out/prof_agents_$P \
    python bench.py --steps 10 --warmup 2 --spinup 40 --no-cpu-baseline --preset $P > gpurun_out/ncu_agents_$P.log 2>&1
tail -2 gpurun_out/ncu_agents_$P.log
