#!/bin/bash
echo "== multi parity quick =="
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "waves_upload or mode_switch" 2>&1 | tail -3
for x in p2p nccl; do
echo "== N=2 exchange=$x =="
SM_EXCHANGE=$x timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 --steps 300 --warmup 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels'])"
done
