#!/bin/bash
# single-GPU vs "fake multi" (strip geometry + multi-GPU agent kernel, no exchange): where does the strip variant lose time?
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
be.step(200)
be.set_timing_enabled(True); be.reset_timing()
be.step(64)
t = be.timing()
print("fake" if fake else "single", "agents_ms", t.agents_ms / t.agent_launches, "trail_ms", t.trail_ms / t.trail_launches, "sort/step", t.sort_ms / 64, "local", be.local_agent_count)
be.close()
PY
SM_FAKE_MULTI=0 python /tmp/fm.py
SM_FAKE_MULTI=1 python /tmp/fm.py
for f in 0 1; do
SM_FAKE_MULTI=$f timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_tex_mem_texture.sum,lts__t_sectors.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none -k regex:'k_agents' -s 230 -c 2 python /tmp/fm.py 2>&1 | grep -E "k_agents|gpu__time|inst_executed|dram__|tex_mem|lts__t|issue_active|l1tex__thr"
done
