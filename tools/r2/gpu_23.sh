#!/bin/bash
# Round 2, GPU call 23 (1 GPU): the racy in-place mode's statistics, the whole GPU suite, smoke, both bench arms.
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
T0=$(date +%s); el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
el "in-place mode"; timeout 300 python -m pytest tests/test_gpu_statistics.py -q -m gpu -x -k inplace 2>&1 | tail -12 | cut -c1-300
el "in-place mode timing"; timeout 100 python - <<'PY' 2>&1 | tail -3
import time, slime_mold_b200 as sm
for flags, name in ((0, "phase_split"), (sm.SM_FLAG_SEM_INPLACE, "in place")):
    be = sm.CudaBackend.new(4096, 4096, sm.Settings.default(), agent_count=16_777_216, flags=flags)
    be.init_agents(1); be.step(100); be.sync(); t0 = time.perf_counter(); be.step(100); be.sync()
    print(name, "us/step", 1e4 * (time.perf_counter() - t0)); be.close()
PY
el "full GPU suite"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2_parity_gpu_c.log
el "smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
el "bench N=1"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_c.log 2> gpurun_out/r2_bench_n1_c.err; tail -1 gpurun_out/r2_bench_n1_c.log | cut -c1-300
el "bench reference"; timeout 400 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tail -1 | cut -c1-600 | tee gpurun_out/r2_bench_ref_n1_c.log
el done
