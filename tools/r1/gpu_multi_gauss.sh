#!/bin/bash
# 2 GPUs: the Gaussian extension's diffusion-only passes on strips (rows kernel at radius 3, streaming kernel at radius 6;
# P2P and NCCL exchange), plus two regular strip cases as a regression check of the exchange path.
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "(gauss or waves_upload or diffuse_mix) and 2-" 2>&1 | tail -15 | tee gpurun_out/r12_multi_gauss.log
