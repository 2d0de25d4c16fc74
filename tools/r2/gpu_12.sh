#!/bin/bash
# Round 2, GPU call 12 (1 GPU): software-pipelined agent kernel -- parity with it forced on, then A/B (same box, same image).
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "parity, pipeline forced on"
SM_AGENT_PIPE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wgsl.py tests/test_gpu_zz_fuzz.py tests/test_gpu_fullsize.py tests/test_gpu_display.py -q -m gpu -x 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2_parity_pipe.log
el "A/B"
for rep in 1 2; do
SM_AGENT_PIPE=0 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_serial | tail -1 | cut -c1-300
SM_AGENT_PIPE=1 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_pipe | tail -1 | cut -c1-300
done
for P in Waves Snake "Firecracker Trees" Curls; do
SM_AGENT_PIPE=0 timeout 120 python tools/probe.py --preset "$P" --steps 48 --spinup 200 --tag "c2_${P}_serial" | tail -1 | cut -c1-300
SM_AGENT_PIPE=1 timeout 120 python tools/probe.py --preset "$P" --steps 48 --spinup 200 --tag "c2_${P}_pipe" | tail -1 | cut -c1-300
done
C3="--agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 24 --spinup 72"
SM_AGENT_PIPE=0 timeout 120 python tools/probe.py $C3 --tag c3_serial | tail -1 | cut -c1-300
SM_AGENT_PIPE=1 timeout 120 python tools/probe.py $C3 --tag c3_pipe | tail -1 | cut -c1-300
SM_AGENT_PIPE=0 timeout 120 python tools/probe.py --agents 1000000 --width 1920 --height 1080 --steps 480 --spinup 480 --tag c1_serial | tail -1 | cut -c1-300
SM_AGENT_PIPE=1 timeout 120 python tools/probe.py --agents 1000000 --width 1920 --height 1080 --steps 480 --spinup 480 --tag c1_pipe | tail -1 | cut -c1-300
SM_AGENT_PIPE=0 timeout 120 python tools/probe.py --steps 48 --spinup 200 --fake-strips 2 --tag c2_strips_serial | tail -1 | cut -c1-300
SM_AGENT_PIPE=1 timeout 120 python tools/probe.py --steps 48 --spinup 200 --fake-strips 2 --tag c2_strips_pipe | tail -1 | cut -c1-300
el "ncu pipe"
SM_AGENT_PIPE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 60 -c 1 -f -o gpurun_out/r2_prof_agents_c2_pipe \
    python tools/probe.py --steps 8 --spinup 80 --no-kernel-split > gpurun_out/r2_ncu_agents_c2_pipe.log 2>&1; tail -1 gpurun_out/r2_ncu_agents_c2_pipe.log
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_pipe.jsonl
el done
