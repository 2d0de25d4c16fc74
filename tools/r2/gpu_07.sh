#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "failing test, verbose"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "ragged" 2>&1 | tail -25 | cut -c1-300
el "same without pairs"
SM_SURF_PAIRS=0 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "ragged" 2>&1 | tail -5 | cut -c1-300
el "same without graphs"
SM_STEP_GRAPH=0 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "ragged" 2>&1 | tail -5 | cut -c1-300
el "A/B pairs"
for rep in 1 2; do
SM_SURF_PAIRS=0 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_rows | tail -1 | cut -c1-300
SM_SURF_PAIRS=1 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_pairs | tail -1 | cut -c1-300
done
SM_STEP_GRAPH=0 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag c2_pairs_nograph | tail -1 | cut -c1-300
SM_STEP_GRAPH=0 timeout 120 python tools/probe.py --agents 1000000 --width 1920 --height 1080 --steps 480 --spinup 480 --tag c1_nograph | tail -1 | cut -c1-300
SM_STEP_GRAPH=1 timeout 120 python tools/probe.py --agents 1000000 --width 1920 --height 1080 --steps 480 --spinup 480 --tag c1_graph | tail -1 | cut -c1-300
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_pairs_ab.jsonl
el done
