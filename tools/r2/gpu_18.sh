#!/bin/bash
# Round 2, GPU call 18 (2 GPUs): the whole multi-GPU parity file (p2p + nccl paths, new ADVICE cases) on the final strip code.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/r2_parity_multi_n2.log
