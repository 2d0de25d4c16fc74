#!/bin/bash
mkdir -p gpurun_out
echo "== fake strips small, plain"; timeout 120 python tools/probe.py --agents 150000 --width 512 --height 192 --steps 35 --spinup 35 --fake-strips 2 --tag small_fake 2>&1 | tail -2 | cut -c1-300
echo "== fake strips small, no boundary first"; SM_BOUNDARY_FIRST=0 timeout 120 python tools/probe.py --agents 150000 --width 512 --height 192 --steps 35 --spinup 35 --fake-strips 2 --tag small_fake 2>&1 | tail -2 | cut -c1-300
echo "== memcheck"; SM_BOUNDARY_FIRST=0 timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python tools/probe.py --agents 150000 --width 512 --height 192 --steps 20 --spinup 20 --fake-strips 2 --no-kernel-split 2>&1 | grep -v "^{" | head -60 | cut -c1-250
