#!/bin/bash
# lean multi-GPU pass: strip parity tests + bench at N = all visible GPUs (P2P exchange)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
echo "== pytest multi =="
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_multi.log
echo "== bench N=$NG =="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $NG --steps 300 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_n${NG}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), d['ms_per_step'], d['kernels'])"
