#!/bin/bash
# multi-GPU: strip parity tests, then bench at N = all GPUs with the overlapped and the serial exchange
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
echo "== pytest multi =="
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_multi.log
for ov in 1 0; do
echo "== bench N=$NG SM_OVERLAP=$ov =="
SM_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2951$ov \
    bench.py --gpus $NG --steps 300 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_n${NG}_ov$ov.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), d['ms_per_step'], d['kernels'], d['diffusion'])"
done
