#!/bin/bash
# Round 2, call 28: k_trail_rows rows per chunk vs whole waves (148 SMs x 8 CTAs = 1184 slots; 4096^2 at 8 rows = 3.46 waves).
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/probe.jsonl
for rep in 1 2; do
for r in 8 10 12 14 16 28 32; do
  SM_TRAIL_ROWS_PER_CHUNK=$r python tools/probe.py --tag c2_rpc$r --steps 96 --spinup 192 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], round(d['us_per_step'],1), round(d['agents_us'],1), round(d['trail_us'],2))"
done
done
for r in 8 14 16 28 32 56; do
  SM_TRAIL_ROWS_PER_CHUNK=$r python tools/probe.py --tag big_rpc$r --agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 32 --spinup 64 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], round(d['us_per_step'],1), round(d['agents_us'],1), round(d['trail_us'],2))"
done
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_trail_rpc.jsonl
