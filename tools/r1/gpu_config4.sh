#!/bin/bash
# BASELINE configs[3]: 1B agents on 32768^2, strips over all visible GPUs; then the weak-scaling bench line at the same N
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
nvidia-smi topo -m > gpurun_out/topo_$NG.txt 2>&1
A=$((1000000000 / NG)); R=$((32768 / NG))
echo "== config 4: $A agents/GPU, 32768 x $R rows/GPU =="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $NG --agents $A --width 32768 --height $R --steps 50 --warmup 5 --spinup 50 2>&1 | tail -3 | tee gpurun_out/bench_config4_n${NG}.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), d['ms_per_step'], d['kernels'], d['diffusion'])"
echo "== weak scaling bench N=$NG =="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $NG --steps 300 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_n${NG}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4e e2e %.4e'%(d['value'], d['e2e']['value']), d['ms_per_step'], d['kernels'])"
