#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
for rep in 1 2; do
timeout 120 python tools/probe.py --steps 48 --spinup 200 --fake-strips 2 --tag strips_5cta | tail -1 | cut -c1-200
SM_LIB_PATH=$PWD/slime_mold_b200/libslime_b200_mb4.so timeout 120 python tools/probe.py --steps 48 --spinup 200 --fake-strips 2 --tag strips_4cta | tail -1 | cut -c1-200
done
timeout 120 python tools/probe.py --steps 48 --spinup 200 --tag single | tail -1 | cut -c1-200
cp gpurun_out/probe.jsonl gpurun_out/r2_probe_strip_ctas.jsonl
