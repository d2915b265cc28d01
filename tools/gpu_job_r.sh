#!/bin/bash
# after the sparse fp64 refinement: parity suite, determinism stress, wall time to links
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2r_gpu_tests.log
timeout 600 python tools/stress_determinism.py 10 2>&1 | tail -4 | tee gpurun_out/r2r_stress.log
timeout 900 python tools/wall_c2.py 2>&1 | tail -3 | cut -c1-1500 | tee gpurun_out/r2r_wall_c2.log
