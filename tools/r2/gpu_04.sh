#!/bin/bash
# Round 2, GPU call 4 (1 GPU): straddle-free dual-copy sampler -- parity, then A/B against the single copy in the same process image.
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "parity: sampler copies, display, presets"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_display.py tests/test_gpu_wgsl.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r2_parity_dual.log
C3="--agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 24 --spinup 72"
for rep in 1 2; do
el "config3 single copy ($rep)";  SM_SAMPLER=tex1 timeout 120 python tools/probe.py $C3 --tag c3_tex1 | tail -1 | cut -c1-420
el "config3 dual copy ($rep)";    SM_SAMPLER=tex2 timeout 120 python tools/probe.py $C3 --tag c3_tex2 | tail -1 | cut -c1-420
done
el "config2 snake single";  SM_SAMPLER=tex1 timeout 120 python tools/probe.py --preset Snake --steps 48 --spinup 200 --tag c2_snake_tex1 | tail -1 | cut -c1-420
el "config2 snake dual";    SM_SAMPLER=tex2 timeout 120 python tools/probe.py --preset Snake --steps 48 --spinup 200 --tag c2_snake_tex2 | tail -1 | cut -c1-420
el "config2 mesh single";   SM_SAMPLER=tex1 timeout 120 python tools/probe.py --preset Mesh --steps 48 --spinup 200 --tag c2_mesh_tex1 | tail -1 | cut -c1-420
el "config2 mesh dual";     SM_SAMPLER=tex2 timeout 120 python tools/probe.py --preset Mesh --steps 48 --spinup 200 --tag c2_mesh_tex2 | tail -1 | cut -c1-420
el "config2 default single"; SM_SAMPLER=tex1 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_default_tex1 | tail -1 | cut -c1-420
el "config2 default dual";   SM_SAMPLER=tex2 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_default_tex2 | tail -1 | cut -c1-420
for sd in 32 48 64 96; do
el "config2 sd $sd single"; SM_SAMPLER=tex1 timeout 120 python tools/probe.py --sd $sd --steps 48 --spinup 120 --tag c2_sd${sd}_tex1 | tail -1 | cut -c1-420
el "config2 sd $sd dual";   SM_SAMPLER=tex2 timeout 120 python tools/probe.py --sd $sd --steps 48 --spinup 120 --tag c2_sd${sd}_tex2 | tail -1 | cut -c1-420
done
el "ncu: k_agents dual, config3"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 30 -c 1 -f -o gpurun_out/r2_prof_agents_sd225_dual \
    python tools/probe.py $C3 --steps 8 --spinup 40 --no-kernel-split > gpurun_out/r2_ncu_agents_sd225_dual.log 2>&1; tail -2 gpurun_out/r2_ncu_agents_sd225_dual.log
el "gauss wring vs stream (R 5-8)"
rm -f gpurun_out/kernel_sweep.jsonl
timeout 200 python tools/bench_kernels.py gauss_packed 2>&1 | tail -2 | cut -c1-300
cp gpurun_out/kernel_sweep.jsonl gpurun_out/r2_gauss_wring_sweep.jsonl 2>/dev/null
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_dual.jsonl
el done
