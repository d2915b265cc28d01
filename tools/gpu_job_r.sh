#!/bin/bash
# after a change of the fp64 refinement / epilogue constants: parity suite, determinism stress, A/B timing, wall time to links
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2r_gpu_tests.log
timeout 600 python tools/stress_determinism.py 10 2>&1 | tail -4 | tee gpurun_out/r2r_stress.log
timeout 900 python tools/kernel_ab.py cur base cur base --nsnp 30000 2>&1 | cut -c1-400 | tee gpurun_out/r2r_ab.log
timeout 900 python tools/wall_c2.py 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/r2r_wall_c2.log
