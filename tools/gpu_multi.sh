#!/bin/bash
# Multi-GPU pass (run with gpurun --gpus N): strip parity tests, then the bench at 1..N GPUs.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== pytest multi =="
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_multi.log
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    echo "== bench N=$n =="
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n$n.log
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 300 --warmup 10 2>&1 | tail -3 | tee gpurun_out/bench_n$n.log
    fi
  fi
done
