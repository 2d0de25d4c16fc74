#!/bin/bash
# Round 2, GPU call 10 (2 GPUs): bench.py at N=2 (parity before timing, config4 block), the multi-GPU parity suite.
mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "bench N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2_b.log 2> gpurun_out/r2_bench_n2_b.err; tail -1 gpurun_out/r2_bench_n2_b.log | cut -c1-600; grep "bench rank 0" gpurun_out/r2_bench_n2_b.err | tail -8
el "multi-GPU parity suite (2 GPUs)"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2_parity_multi_n2.log
el done
