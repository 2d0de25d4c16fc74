#!/bin/bash
# Round 2, GPU call 6: whole-sector surface writes in k_trail_rows -- parity, then timing (single and dual copy).
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "parity"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_display.py tests/test_gpu_wgsl.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r2_parity_sust_pairs.log
C3="--agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 24 --spinup 72"
el "config3 single";  SM_SAMPLER=tex1 timeout 120 python tools/probe.py $C3 --tag c3_tex1 | tail -1 | cut -c1-420
el "config3 dual";    SM_SAMPLER=tex2 timeout 120 python tools/probe.py $C3 --tag c3_tex2 | tail -1 | cut -c1-420
el "config2 snake single";  SM_SAMPLER=tex1 timeout 120 python tools/probe.py --preset Snake --steps 48 --spinup 200 --tag c2_snake_tex1 | tail -1 | cut -c1-420
el "config2 snake dual";    SM_SAMPLER=tex2 timeout 120 python tools/probe.py --preset Snake --steps 48 --spinup 200 --tag c2_snake_tex2 | tail -1 | cut -c1-420
el "config2 default single"; SM_SAMPLER=tex1 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_default_tex1 | tail -1 | cut -c1-420
el "config2 default dual";   SM_SAMPLER=tex2 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_default_tex2 | tail -1 | cut -c1-420
el "config2 dep 0.3";        timeout 120 python tools/probe.py --steps 48 --spinup 200 --dep 0.3 --tag c2_dep03 | tail -1 | cut -c1-420
el "config1";  timeout 120 python tools/probe.py --agents 1000000 --width 1920 --height 1080 --steps 480 --spinup 480 --tag c1 | tail -1 | cut -c1-420
el "ncu trail 4096"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_trail_rows' -s 30 -c 1 -f -o gpurun_out/r2_prof_trail_4096_pairs \
    python tools/probe.py --steps 8 --spinup 40 --no-kernel-split > gpurun_out/r2_ncu_trail_4096_pairs.log 2>&1; tail -1 gpurun_out/r2_ncu_trail_4096_pairs.log
el "launch list config1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 60 --csv --log-file gpurun_out/r2_launches_config1.csv \
    python tools/probe.py --agents 1000000 --width 1920 --height 1080 --steps 100 --spinup 300 --no-kernel-split > /dev/null 2>&1
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_sust_pairs.jsonl
el done
