#!/bin/bash
# Multi-GPU pass (run with gpurun --gpus N): single-GPU parity, strip parity tests, then the bench at 1..N GPUs.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== pytest single-GPU parity =="
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_math.py -x -q -m gpu 2>&1 | tail -3
echo "== pytest multi =="
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_multi.log
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    echo "== bench N=$n =="
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 300 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n$n.log
    else
      for x in p2p nccl; do SM_EXCHANGE=$x timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 300 --warmup 10 2>&1 | tail -1 | tee gpurun_out/bench_n${n}_$x.log; done; cp gpurun_out/bench_n${n}_p2p.log gpurun_out/bench_n$n.log
    fi
  fi
done
python - <<'PY'
import json
import glob
for f in sorted(glob.glob("gpurun_out/bench_n*.log")):
    n = f
    try:
        for l in open(f):
            if l.startswith("{"):
                d = json.loads(l)
                print(n, "value %.3e e2e %.3e ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["kernels"])
    except FileNotFoundError:
        pass
PY
