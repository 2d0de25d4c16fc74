#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/fm.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import slime_mold_b200 as sm
W, H, N = 4096, 8192, 33554432
fake = os.environ.get("SM_FAKE_MULTI") == "1"
if fake:
    be = sm.CudaBackend.new(W, H, agent_count=N, rank=0, world_size=2)
else:
    be = sm.CudaBackend.new(W, H // 2, agent_count=N // 2)
be.init_agents(1)
be.step(120)
be.set_timing_enabled(True); be.reset_timing()
be.step(64)
t = be.timing()
print("fake" if fake else "single", "agents_ms", t.agents_ms / t.agent_launches, "trail_ms", t.trail_ms / t.trail_launches, "sort/step", t.sort_ms / 64)
be.close()
PY
SM_FAKE_MULTI=0 python /tmp/fm.py
SM_FAKE_MULTI=1 python /tmp/fm.py
SM_FAKE_MULTI=1 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_agents' -s 150 -c 2 -f -o gpurun_out/prof_fake_multi python /tmp/fm.py > gpurun_out/ncu_fake.log 2>&1
SM_FAKE_MULTI=0 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_agents' -s 150 -c 2 -f -o gpurun_out/prof_fake_single python /tmp/fm.py > gpurun_out/ncu_fake_single.log 2>&1
