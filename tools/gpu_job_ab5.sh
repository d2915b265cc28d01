#!/bin/bash
mkdir -p gpurun_out
V="$@"
timeout 900 python tools/kernel_ab.py head $V head $V --nsnp 30000 2>&1 | cut -c1-100 | tee gpurun_out/ab5.log
for v in $V; do echo "== stress $v"; LDW_LIBRARY_PATH=$PWD/ldweaver_b200/variants/libldwgpu_$v.so timeout 900 python tools/stress_determinism.py 30 2>&1 | grep -v "run [0-9]*:" | tail -5; done | tee gpurun_out/ab5_stress.log
