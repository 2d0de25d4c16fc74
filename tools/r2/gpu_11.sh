#!/bin/bash
# Round 2, GPU call 11 (1 GPU): after the sm_tuning refactor -- whole suite, smoke, bench N=1, the strip kernels in one process.
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "full GPU suite"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_parity_gpu_b.log
el "statistics table"; timeout 300 python -m pytest tests/test_gpu_statistics.py -q -m gpu -s 2>&1 | tail -25 | cut -c1-300 | tee gpurun_out/r2_statistics_config1.log
el "smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
el "bench N=1"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_b.log 2> gpurun_out/r2_bench_n1_b.err; tail -1 gpurun_out/r2_bench_n1_b.log | cut -c1-400; tail -3 gpurun_out/r2_bench_n1_b.err
el "strip kernels, one process (self-peer): config 2 per rank"
timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_single | tail -1 | cut -c1-400
timeout 120 python tools/probe.py --steps 48 --spinup 200 --fake-strips 2 --tag c2_fake_strips | tail -1 | cut -c1-400
SM_SIDE_TIMING=1 timeout 120 python tools/probe.py --steps 48 --spinup 200 --fake-strips 2 --tag c2_fake_strips_side 2>&1 | tail -3 | cut -c1-400
el "ncu: k_agents<XM_P2P> (self-peer)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 60 -c 1 -f -o gpurun_out/r2_prof_agents_p2p \
    python tools/probe.py --steps 8 --spinup 80 --fake-strips 2 --no-kernel-split > gpurun_out/r2_ncu_agents_p2p.log 2>&1; tail -1 gpurun_out/r2_ncu_agents_p2p.log
el "ncu: k_agents single, config 2 (source counters)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_agents' -s 60 -c 1 -f -o gpurun_out/r2_prof_agents_c2 \
    python tools/probe.py --steps 8 --spinup 80 --no-kernel-split > gpurun_out/r2_ncu_agents_c2.log 2>&1; tail -1 gpurun_out/r2_ncu_agents_c2.log
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_strips.jsonl
el done
