#!/bin/bash
# Round 2, GPU call 17 (1 GPU): maps beyond 2^31 cells, 65536^2 diffusion, statistics, warp-aggregated counts A/B.
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
free -g | head -2; nproc
el "big maps"; timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -x -k "beyond or 65536" 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r2_parity_bigmaps.log
el "statistics"; timeout 300 python -m pytest tests/test_gpu_statistics.py -q -m gpu -s 2>&1 | tail -24 | cut -c1-250 | tee gpurun_out/r2_statistics_config1.log
el "match_any counts: parity (dep < 1 cases)"
SM_LIB_PATH=$PWD/slime_mold_b200/libslime_b200_agg.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wgsl.py tests/test_gpu_zz_fuzz.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
el "match_any counts: A/B at dep 0.3"
for rep in 1 2; do
timeout 120 python tools/probe.py --steps 48 --spinup 200 --dep 0.3 --tag c2_dep03_red | tail -1 | cut -c1-200
SM_LIB_PATH=$PWD/slime_mold_b200/libslime_b200_agg.so timeout 120 python tools/probe.py --steps 48 --spinup 200 --dep 0.3 --tag c2_dep03_match_any | tail -1 | cut -c1-200
done
C3="--agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 24 --spinup 72"
timeout 120 python tools/probe.py $C3 --dep 0.3 --tag c3_dep03_red | tail -1 | cut -c1-200
SM_LIB_PATH=$PWD/slime_mold_b200/libslime_b200_agg.so timeout 120 python tools/probe.py $C3 --dep 0.3 --tag c3_dep03_match_any | tail -1 | cut -c1-200
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_match_any.jsonl
el done
