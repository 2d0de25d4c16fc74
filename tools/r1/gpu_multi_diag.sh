#!/bin/bash
NG=$(nvidia-smi -L | wc -l)
SM_SIDE_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $NG --steps 300 --warmup 10 2>&1 | grep -E "side stream|^\{|\[rank" | sed -E 's/.*("value": [0-9.]+).*("ms_per_step": [0-9.]+).*("diffusion": \{[^}]*\}).*/\1 \2 \3/' | cut -c1-500
