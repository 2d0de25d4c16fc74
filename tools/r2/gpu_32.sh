#!/bin/bash
# Round 2, call 32: u8 deposit flags in 8x8 tiles (default) vs row-major (SM_FLAG_LAYOUT=linear): parity, then A/B in one call.
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_display.py tests/test_gpu_wgsl.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_parity_tiled_flags.log
grep -q "failed\|error" gpurun_out/r2_parity_tiled_flags.log && exit 1
rm -f gpurun_out/probe.jsonl
show='import sys,json; d=json.loads(sys.stdin.read()); print(d["tag"], round(d["us_per_step"],1), round(d["agents_us"],1), round(d["trail_us"],2), round(d["sort_us_per_step"],1))'
for rep in 1 2; do
for lin in 1 0; do
  SM_FLAG_LAYOUT=$( [ $lin = 1 ] && echo linear || echo tiled ) python tools/probe.py --tag big_linear$lin --agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 96 --spinup 96 2>&1 | tail -1 | python -c "$show"
  SM_FLAG_LAYOUT=$( [ $lin = 1 ] && echo linear || echo tiled ) python tools/probe.py --tag c2_linear$lin --steps 96 --spinup 192 2>&1 | tail -1 | python -c "$show"
done
done
for lin in 1 0; do
  SM_FLAG_LAYOUT=$( [ $lin = 1 ] && echo linear || echo tiled ) python tools/probe.py --tag c1_linear$lin --agents 1000000 --width 1920 --height 1080 --steps 960 --spinup 960 2>&1 | tail -1 | python -c "$show"
  SM_FLAG_LAYOUT=$( [ $lin = 1 ] && echo linear || echo tiled ) python tools/probe.py --tag snake_linear$lin --preset Snake --steps 96 --spinup 192 2>&1 | tail -1 | python -c "$show"
done
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_tiled_flags.jsonl
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "config2" 2>&1 | tail -3
