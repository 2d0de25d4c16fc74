#!/bin/bash
# Round 2, GPU call 1 (unchanged round-1 build): where does k_agents spend its time at sensor distance 225 (BASELINE configs[2])?
# A/B of the experiments prepared in round 1 (sort keys, stream hint, sampler, L2 fetch granularity), ncu captures, experiment parity.
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
C3="--agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 24 --spinup 72"
el "config3 default";            timeout 120 python tools/probe.py $C3 --tag c3_default | tail -1
el "config3 snake preset";       timeout 120 python tools/probe.py --agents 100000000 --width 8192 --height 8192 --preset Snake --steps 24 --spinup 72 --tag c3_snake | tail -1
el "config3 stream hint";        SM_AGENT_STREAM_HINT=1 timeout 120 python tools/probe.py $C3 --tag c3_hint | tail -1
el "config3 ldg sampler";        SM_SAMPLER=ldg timeout 120 python tools/probe.py $C3 --tag c3_ldg | tail -1
el "config3 l2 gran 32";         SM_L2_FETCH_GRANULARITY=32 timeout 120 python tools/probe.py $C3 --tag c3_gran32 | tail -1
el "config3 heading bins 16";    SM_SORT_HEADING_BINS=16 timeout 120 python tools/probe.py $C3 --tag c3_hb16 | tail -1
el "config3 super shift 2";      SM_SORT_SUPER_SHIFT=2 timeout 120 python tools/probe.py $C3 --tag c3_ss2 | tail -1
el "config3 tiles 16x16";        SM_TILE_SHIFT_X=4 SM_TILE_SHIFT_Y=4 timeout 120 python tools/probe.py $C3 --tag c3_tile16 | tail -1
el "config3 tiles 32x4";         SM_TILE_SHIFT_X=5 SM_TILE_SHIFT_Y=2 timeout 120 python tools/probe.py $C3 --tag c3_tile32x4 | tail -1
el "config3 sort every 8";       timeout 120 python tools/probe.py $C3 --sort-interval 8 --tag c3_sort8 | tail -1
el "config2 default";            timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_default | tail -1
el "config2 super shift 2";      SM_SORT_SUPER_SHIFT=2 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_ss2 | tail -1
el "config2 dep 0.3 (counts)";   timeout 120 python tools/probe.py --steps 48 --spinup 200 --dep 0.3 --tag c2_dep03 | tail -1
el "config2 snake";              timeout 120 python tools/probe.py --preset Snake --steps 48 --spinup 200 --tag c2_snake | tail -1
el "ncu full: k_agents config3"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 30 -c 1 -f -o gpurun_out/r2_prof_agents_sd225 \
    python tools/probe.py $C3 --steps 8 --spinup 40 --no-kernel-split > gpurun_out/r2_ncu_agents_sd225.log 2>&1; tail -2 gpurun_out/r2_ncu_agents_sd225.log
el "ncu full: k_trail_rows config3"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_trail_rows' -s 30 -c 1 -f -o gpurun_out/r2_prof_trail_8192 \
    python tools/probe.py $C3 --steps 8 --spinup 40 --no-kernel-split > gpurun_out/r2_ncu_trail_8192.log 2>&1; tail -2 gpurun_out/r2_ncu_trail_8192.log
el "experiment parity"
SM_TEST_EXPERIMENTS=1 timeout 300 python -m pytest tests/test_gpu_zz_fuzz.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2_parity_experiments.log
el done
