#!/bin/bash
# Round 2, call 33: sort interval again, now that the deposit flags are tiled (the agent kernel loses less as the order decays).
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/probe.jsonl
show='import sys,json; d=json.loads(sys.stdin.read()); print(d["tag"], round(d["us_per_step"],1), round(d["agents_us"],1), round(d["trail_us"],2), round(d["sort_us_per_step"],1))'
for si in 24 16 32 40 48; do
  python tools/probe.py --tag big_si$si --sort-interval $si --agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 480 --spinup 96 2>&1 | tail -1 | python -c "$show"
done
for si in 24 32 48; do
  python tools/probe.py --tag c2_si$si --sort-interval $si --steps 480 --spinup 192 2>&1 | tail -1 | python -c "$show"
done
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_sort_interval_tiled.jsonl
