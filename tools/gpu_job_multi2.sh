#!/bin/bash
# 2 GPUs: device-group tests + bench C2 under torchrun
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2u_multi_tests_2gpu.log
bash tools/gpu_job_scale.sh 2 C2 10
