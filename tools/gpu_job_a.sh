#!/bin/bash
# round-2 GPU job A: parity suite, sanitizer runs on the smoke, ncu --set full for the non-scan kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_gpu_tests.log
tail -5 gpurun_out/r2a_gpu_tests.log
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2a_memcheck.log python __graft_entry__.py smoke > gpurun_out/r2a_memcheck.out 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r2a_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/r2a_racecheck.log python __graft_entry__.py smoke > gpurun_out/r2a_racecheck.out 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r2a_racecheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hdw_gemm_kernel|column_count_kernel|extract_codes_kernel' -c 6 -o gpurun_out/r2a_stage_kernels python tools/bench_stages.py > gpurun_out/r2a_ncu_stage.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -8
