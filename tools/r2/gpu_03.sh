#!/bin/bash
# Round 2, GPU call 3 (2 GPUs): the new bench.py contract at N=1 and N=2, both arms.
mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nproc > gpurun_out/nproc.txt
el "bench N=1"; timeout 600 python bench.py > gpurun_out/r2_bench_n1_a.log 2> gpurun_out/r2_bench_n1_a.err; tail -1 gpurun_out/r2_bench_n1_a.log | cut -c1-3000; tail -3 gpurun_out/r2_bench_n1_a.err
el "bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2_a.log 2> gpurun_out/r2_bench_n2_a.err; tail -1 gpurun_out/r2_bench_n2_a.log | cut -c1-3000; tail -5 gpurun_out/r2_bench_n2_a.err
el "reference N=1"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 --ref-seconds 60 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/r2_bench_ref_n1_a.log
el done
