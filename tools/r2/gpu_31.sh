#!/bin/bash
# Round 2, call 31: what does the deposit flag store cost k_agents, and would a tiled flag layout make it cheaper?
# A/B builds that issue a SECOND flag store into a dummy buffer: dup1 linear (same address pattern as the real one),
# dup2 8x8-cell tiles of 64 B, dup3 16x8-cell tiles of 128 B.  (dup1 - base) = cost of the linear store; (dupK - base) = cost of a tiled one.
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/probe.jsonl
show='import sys,json; d=json.loads(sys.stdin.read()); print(d["tag"], round(d["us_per_step"],1), round(d["agents_us"],1), round(d["trail_us"],2), round(d["sort_us_per_step"],1))'
for rep in 1 2; do
for v in base dup1 dup2 dup3; do
  SM_LIB_PATH=$PWD/ab/$v.so python tools/probe.py --tag big_$v --agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 96 --spinup 96 2>&1 | tail -1 | python -c "$show"
done
done
for v in base dup1 dup2 dup3; do
  SM_LIB_PATH=$PWD/ab/$v.so python tools/probe.py --tag c2_$v --steps 96 --spinup 192 2>&1 | tail -1 | python -c "$show"
done
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_deposit_layout.jsonl
