#!/bin/bash
# Round 2, call 34 (final tree, 1 GPU): the whole GPU suite, then ncu captures of the two step kernels with the tiled flag field.
cd "$GRAFT_REPO_ROOT"
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_parity_gpu_final.log
el "suite done"
C3="--agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 30 -c 1 -f -o gpurun_out/r2_prof_agents_sd225_tiled \
    python tools/probe.py $C3 --steps 8 --spinup 40 --no-kernel-split > gpurun_out/r2_ncu_agents_sd225_tiled.log 2>&1; tail -1 gpurun_out/r2_ncu_agents_sd225_tiled.log | cut -c1-200
el "ncu agents done"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_trail_rows' -s 30 -c 1 -f -o gpurun_out/r2_prof_trail_8192_tiled \
    python tools/probe.py $C3 --steps 8 --spinup 40 --no-kernel-split > gpurun_out/r2_ncu_trail_8192_tiled.log 2>&1; tail -1 gpurun_out/r2_ncu_trail_8192_tiled.log | cut -c1-200
el "ncu trail done"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 60 -c 1 -f -o gpurun_out/r2_prof_agents_c2_tiled \
    python tools/probe.py --steps 8 --spinup 80 --no-kernel-split > gpurun_out/r2_ncu_agents_c2_tiled.log 2>&1; tail -1 gpurun_out/r2_ncu_agents_c2_tiled.log | cut -c1-200
el done
