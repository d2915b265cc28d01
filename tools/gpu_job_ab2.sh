#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mi.py -x -q 2>&1 | tail -25 | tee gpurun_out/ab2_tests.log
timeout 900 python tools/kernel_ab.py head base head base --nsnp 30000 2>&1 | tee gpurun_out/ab2.log
LDW_DBG_BLOCK=4 timeout 300 python tools/kernel_ab.py --worker base --data /tmp/kernel_ab_data.npz --steps 1 2>&1 | grep -i "dbg\|AB" | cut -c1-600 | tail -3 | tee gpurun_out/ab2_dbg.log
