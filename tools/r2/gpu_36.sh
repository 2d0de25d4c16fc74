#!/bin/bash
# Round 2, call 36 (1 GPU): tiled flags generalised to strips -- single-GPU parity again, then the strip kernels in the
# self-peer profiling mode under compute-sanitizer (addresses only; results are meaningless in that mode).
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_display.py -m gpu -x -q 2>&1 | tail -4
SM_FLAG_LAYOUT=tiled timeout 300 compute-sanitizer --tool memcheck python tools/probe.py --fake-strips 2 --agents 200000 --width 512 --height 256 --steps 30 --spinup 60 --tag fake_tiled 2>&1 | tail -6 | cut -c1-300
SM_FLAG_LAYOUT=tiled SM_OVERLAP=0 timeout 300 compute-sanitizer --tool memcheck python tools/probe.py --fake-strips 2 --agents 200000 --width 512 --height 256 --steps 30 --spinup 60 --tag fake_tiled_serial 2>&1 | tail -4 | cut -c1-300
