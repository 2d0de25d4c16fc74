#!/bin/bash
# Round 2, GPU call 21 (2 GPUs): Gaussian full steps and the display pass on strips, the capacity fix (empty_strip), partial uploads.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "gauss_rows_full or gauss_stream_full or render or empty_strip or partial_upload or default_devinit" 2>&1 | grep -vE "^\s*$" | tail -30 | cut -c1-300 | tee gpurun_out/r2_parity_multi_n2_new.log
