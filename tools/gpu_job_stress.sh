#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"
  if [ "$v" = "base" ]; then timeout 900 python tools/stress_determinism.py ${REPS:-20} 2>&1 | tail -8
  else LDW_LIBRARY_PATH=$PWD/ldweaver_b200/variants/libldwgpu_$v.so timeout 900 python tools/stress_determinism.py ${REPS:-20} 2>&1 | tail -8; fi
done | tee gpurun_out/stress.log
