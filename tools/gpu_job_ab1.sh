#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python tools/kernel_ab.py head pf rows rows_rcp3 head --nsnp 30000 2>&1 | tee gpurun_out/ab1.log
timeout 900 python -m pytest tests/test_gpu_mi.py -x -q 2>&1 | tail -15 | tee gpurun_out/ab1_tests.log
LDW_DBG_BLOCK=3 LDW_LIBRARY_PATH=$PWD/ldweaver_b200/variants/libldwgpu_rows.so timeout 300 python tools/kernel_ab.py --worker rows --data /tmp/kernel_ab_data.npz --steps 1 2>&1 | grep -i "dbg\|AB" | cut -c1-1200 | tee gpurun_out/ab1_dbg.log
