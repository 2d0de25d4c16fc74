#!/bin/bash
# Round 2, GPU call 19 (1 GPU): the unrolled agent kernel -- parity, A/B against the rolled one.
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wgsl.py tests/test_gpu_zz_fuzz.py tests/test_gpu_fullsize.py -q -m gpu -x -k "not beyond and not 65536" 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2_parity_u4.log
el "A/B"
for rep in 1 2; do
SM_AGENT_ROLLED=1 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_rolled | tail -1 | cut -c1-200
SM_AGENT_ROLLED=0 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_unrolled | tail -1 | cut -c1-200
done
for P in Waves Snake; do
SM_AGENT_ROLLED=1 timeout 120 python tools/probe.py --preset "$P" --steps 48 --spinup 200 --tag "c2_${P}_rolled" | tail -1 | cut -c1-200
SM_AGENT_ROLLED=0 timeout 120 python tools/probe.py --preset "$P" --steps 48 --spinup 200 --tag "c2_${P}_unrolled" | tail -1 | cut -c1-200
done
C3="--agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 24 --spinup 72"
SM_AGENT_ROLLED=1 timeout 120 python tools/probe.py $C3 --tag c3_rolled | tail -1 | cut -c1-200
SM_AGENT_ROLLED=0 timeout 120 python tools/probe.py $C3 --tag c3_unrolled | tail -1 | cut -c1-200
el "ncu unrolled"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 60 -c 1 -f -o gpurun_out/r2_prof_agents_c2_u4 \
    python tools/probe.py --steps 8 --spinup 80 --no-kernel-split > gpurun_out/r2_ncu_agents_c2_u4.log 2>&1; tail -1 gpurun_out/r2_ncu_agents_c2_u4.log
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_u4.jsonl
el done
