#!/bin/bash
# Round 2, call 30: sort interval at the headline workload (the gathers at sensor distance 225 do not profit from the cell
# order; deposits and the step's own cost of sorting do) and at configs[1].
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/probe.jsonl
show='import sys,json; d=json.loads(sys.stdin.read()); print(d["tag"], round(d["us_per_step"],1), round(d["agents_us"],1), round(d["trail_us"],2), round(d["sort_us_per_step"],1))'
for si in 24 12 36 48 72 96; do
  python tools/probe.py --tag big_si$si --sort-interval $si --agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 288 --spinup 96 2>&1 | tail -1 | python -c "$show"
done
for si in 24 32 48; do
  python tools/probe.py --tag c2_si$si --sort-interval $si --steps 288 --spinup 192 2>&1 | tail -1 | python -c "$show"
done
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_sort_interval.jsonl
