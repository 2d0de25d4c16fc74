#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "full GPU suite"
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r2_parity_gpu_a.log
el "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
el done
