#!/bin/bash
# strip parity (P2P path) + bench at N = all GPUs with side-stream timing
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "p2p and ${NG}-" 2>&1 | tail -4 | tee gpurun_out/pytest_multi_p2p_$NG.log
SM_SIDE_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $NG --steps 300 --warmup 10 2>&1 | grep -E "side stream|^\{|\[rank" | tee gpurun_out/bench_n${NG}_diag.log | sed -E 's/.*("value": [0-9.]+).*("ms_per_step": [0-9.]+).*("diffusion": \{[^}]*\}).*/\1 \2 \3/' | cut -c1-420
grep -E "^\{" gpurun_out/bench_n${NG}_diag.log > gpurun_out/bench_n${NG}.log
