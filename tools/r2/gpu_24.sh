#!/bin/bash
# Round 2, GPU call 24 (8 GPUs): bench.py --gpus 8 again (the diffusion block now fits the 65536-row limit).
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/r2_bench_n8.log 2> gpurun_out/r2_bench_n8.err; tail -1 gpurun_out/r2_bench_n8.log | cut -c1-300; grep "bench rank 0" gpurun_out/r2_bench_n8.err | tail -6
