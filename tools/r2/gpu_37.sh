#!/bin/bash
# Round 2, call 37 (2 GPUs): tiled deposit flags on strips -- the tiled_* strip cases against the oracle, then the 2-GPU headline
# with row-major and tiled flags in the same call.
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "tiled_" 2>&1 | tail -8 | tee gpurun_out/r2_parity_multi_tiled.log
for lay in linear auto; do
SM_FLAG_LAYOUT=$lay python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 100 --warmup 5 --only-headline > gpurun_out/r2_bench_n2_$lay.log 2> gpurun_out/r2_bench_n2_$lay.err
python - $lay <<'PY'
import json, sys
for l in open(f'gpurun_out/r2_bench_n2_{sys.argv[1]}.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print(sys.argv[1], d['value'], d['ms_per_step'], d['parity_n']['agents_equal'], d['parity_n']['trail_equal'], d['kernels']['agents']['ms'], d['kernels']['trail']['ms'], d['kernels']['exchange_ms_per_step'])
PY
done
