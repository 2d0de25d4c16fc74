#!/bin/bash
# Round 2, call 29 (2 GPUs): sm_resize on strips; NVLink byte counters around a headline run; then the whole strip suite.
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "resize" 2>&1 | tail -15 | tee gpurun_out/r2_parity_multi_resize.log
grep -q "failed\|error" gpurun_out/r2_parity_multi_resize.log && exit 1
nvidia-smi nvlink -gt d > gpurun_out/r2_nvlink_before.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 200 --warmup 5 --only-headline > gpurun_out/r2_bench_n2_nvlink.log 2> gpurun_out/r2_bench_n2_nvlink.err
nvidia-smi nvlink -gt d > gpurun_out/r2_nvlink_after.txt 2>&1
tail -c 600 gpurun_out/r2_bench_n2_nvlink.log
head -12 gpurun_out/r2_nvlink_after.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2_parity_multi_n2_final.log
