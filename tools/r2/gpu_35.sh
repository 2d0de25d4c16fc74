#!/bin/bash
# Round 2, call 35 (2 GPUs, final tree): a cross-section of the strip cases on both exchange paths, then the 2-GPU headline.
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "waves_upload or default_devinit or resize or render or gauss_rows_full or mode_switch or empty_strip" 2>&1 | tail -5 | tee gpurun_out/r2_parity_multi_n2_final2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 5 --no-config4 > gpurun_out/r2_bench_n2_e.log 2> gpurun_out/r2_bench_n2_e.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_n2_e.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['value'], d['ms_per_step'], d['parity_n'], d['config2_default']['value'], d['config2_default']['ms_per_step'], d['clocks']['reasons'])
PY
tail -2 gpurun_out/r2_bench_n2_e.err
