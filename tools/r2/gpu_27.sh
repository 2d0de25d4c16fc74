#!/bin/bash
# Round 2, call 27: k_trail_rows addressing diet -- A/B of three builds in one call (tools/build_ab.sh):
#   head8 = per-load 64-bit multiplies (previous commit), lean8 = one running cell index, 64 registers (spills 48 B),
#   lean7 = the same at 7 CTAs per SM (71 registers, no spills).  Then the parity tests on the in-tree build (lean8).
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/probe.jsonl
for rep in 1 2; do
for v in head8 lean8 lean7; do
  SM_LIB_PATH=$PWD/ab/$v.so python tools/probe.py --tag c2_$v --steps 96 --spinup 192 2>&1 | tail -1 | cut -c1-400
done
done
for v in head8 lean8 lean7; do
  SM_LIB_PATH=$PWD/ab/$v.so python tools/probe.py --tag big_$v --agents 100000000 --width 8192 --height 8192 --sd 225 --sa 1.34 --steps 32 --spinup 64 2>&1 | tail -1 | cut -c1-400
  SM_LIB_PATH=$PWD/ab/$v.so python tools/probe.py --tag counts_$v --dep 0.5 --steps 48 --spinup 96 2>&1 | tail -1 | cut -c1-400
done
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_trail_addressing.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
