#!/bin/bash
# A/B: diffusion-only kernel of an older build vs the current one, same box, back to back, twice.
for rep in 1 2; do
for lib in slime_mold_b200/libslime_b200_old.so slime_mold_b200/libslime_b200.so; do
  echo "== $lib =="
  SM_LIB_PATH=$PWD/$lib python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import slime_mold_b200 as sm
for S in (4096, 16384):
    be = sm.CudaBackend.new(S, S, agent_count=1)
    be.write_trail(np.random.default_rng(0).random((256, S), dtype=np.float32))
    be.diffuse_only(10); be.sync()
    st = torch.cuda.ExternalStream(be.stream_handle)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200 if S == 4096 else 40
    e0.record(st); be.diffuse_only(n); e1.record(st); e1.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(S, "ms/pass %.4f GB/s %.0f" % (ms, 8.0 * S * S / ms / 1e6))
    be.close()
PY
done
done
nvidia-smi --query-gpu=clocks.mem,clocks.max.mem,clocks.sm,power.draw,temperature.gpu --format=csv
