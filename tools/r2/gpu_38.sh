#!/bin/bash
# Round 2, call 38 (4 GPUs): tiled deposit flags on strips with two DISTINCT ring neighbours per rank: strip cases against the
# oracle at world size 4, then the 4-GPU headline (its parity block now runs both flag layouts before timing).
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "tiled_default_devinit or tiled_waves_upload or tiled_thin_strips or tiled_render or tiled_empty_strip" 2>&1 | tail -5 | tee gpurun_out/r2_parity_multi_tiled_n4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 50 --warmup 3 --only-headline > gpurun_out/r2_bench_n4_tiled.log 2> gpurun_out/r2_bench_n4_tiled.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_n4_tiled.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['value'], d['ms_per_step'], d['parity_n']['agents_equal'], d['parity_n']['trail_equal'], d['kernels']['agents']['ms'], d['kernels']['trail']['ms'], d['kernels']['exchange_ms_per_step'], d['clocks']['reasons'])
PY
tail -2 gpurun_out/r2_bench_n4_tiled.err
