#!/bin/bash
# Round 2, GPU call 22 (8 GPUs): bench.py --gpus 8 (headline weak-scaled configs[2], config2_default, config4 = 1 B agents on 32768^2,
# diffusion, parity before timing), then strip parity cases at world 4 and 8.
mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi -L | head -8
el "bench N=8"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/r2_bench_n8.log 2> gpurun_out/r2_bench_n8.err; tail -1 gpurun_out/r2_bench_n8.log | cut -c1-400; grep "bench rank 0" gpurun_out/r2_bench_n8.err | tail -6
el "strip parity at world 4 / 8"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "8-default_devinit-p2p or 8-firecracker_devinit-p2p or 8-render-p2p or 8-gauss_rows_full-p2p or 4-waves_upload-p2p or 4-diffuse_mix-p2p or 4-empty_strip-p2p or 8-mode_switch-p2p" 2>&1 | grep -vE "^\s*$" | tail -12 | cut -c1-300 | tee gpurun_out/r2_parity_multi_n8.log
el "bench N=4"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 50 --warmup 3 > gpurun_out/r2_bench_n4.log 2> gpurun_out/r2_bench_n4.err; tail -1 gpurun_out/r2_bench_n4.log | cut -c1-300
el done
