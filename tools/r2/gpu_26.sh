#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
C3="--agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 24 --spinup 72"
for shape in "3 3" "3 4" "2 4" "2 3" "3 5"; do
set -- $shape
SM_TILE_SHIFT_X=$1 SM_TILE_SHIFT_Y=$2 timeout 120 python tools/probe.py $C3 --tag "c3_tile_$1_$2" | tail -1 | cut -c1-120
done
for shape in "3 4" "2 3"; do
set -- $shape
SM_TILE_SHIFT_X=$1 SM_TILE_SHIFT_Y=$2 timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag "c2_tile_$1_$2" | tail -1 | cut -c1-120
done
cat gpurun_out/probe.jsonl >> gpurun_out/r2_probe_tiles.jsonl
