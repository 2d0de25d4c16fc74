#!/bin/bash
# Round 2, GPU call 15 (2 GPUs): boundary-first stepping on strips -- multi-GPU parity, then bench N=2 with and without it.
mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "multi-GPU parity suite (2 GPUs), p2p cases"
timeout 500 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "p2p" 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2_parity_multi_n2_split.log
el "bench N=2, boundary first"
SM_SIDE_TIMING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 96 --warmup 3 --no-config4 > gpurun_out/r2_bench_n2_c.log 2> gpurun_out/r2_bench_n2_c.err; grep "side stream" gpurun_out/r2_bench_n2_c.err | tail -4
el "bench N=2, single agent launch"
SM_BOUNDARY_FIRST=0 SM_SIDE_TIMING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 96 --warmup 3 --no-config4 > gpurun_out/r2_bench_n2_d.log 2> gpurun_out/r2_bench_n2_d.err; grep "side stream" gpurun_out/r2_bench_n2_d.err | tail -4
el done
