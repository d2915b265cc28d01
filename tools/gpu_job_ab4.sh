#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kernel_ab.py head rows2 r104 head rows2 r104 --nsnp 30000 2>&1 | tee gpurun_out/ab4.log
LDW_DBG_BLOCK=4 LDW_LIBRARY_PATH=$PWD/ldweaver_b200/variants/libldwgpu_r104.so timeout 300 python tools/kernel_ab.py --worker base --data /tmp/kernel_ab_data.npz --steps 1 2>&1 | grep -i "dbg" | cut -c1-600 | tail -1 | tee gpurun_out/ab4_dbg.log
