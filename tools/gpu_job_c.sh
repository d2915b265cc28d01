#!/bin/bash
# round-2 GPU job C: full-size parity (C3, C4, C5), bench lines of every BASELINE config on one GPU, ncu of the HDW GEMM
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k "c4_full or c5_full or c3_full" ) > gpurun_out/r2c_fullsize_tests.log 2>&1
tail -8 gpurun_out/r2c_fullsize_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err; echo "C2 rc=$?"; tail -c 600 gpurun_out/r2c_bench_c2.err
timeout 600 python bench.py --config C3 --steps 5 --warmup 2 > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err; echo "C3 rc=$?"; tail -c 300 gpurun_out/r2c_bench_c3.err
timeout 1200 python bench.py --config C4 --steps 3 --warmup 1 --no-cpu > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err; echo "C4 rc=$?"; tail -c 600 gpurun_out/r2c_bench_c4.err
timeout 1200 python bench.py --config C5 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err; echo "C5 rc=$?"; tail -c 600 gpurun_out/r2c_bench_c5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hdw_gemm_kernel' -c 2 -o gpurun_out/r2c_hdw_gemm python bench.py --config C3 --steps 1 --warmup 0 > gpurun_out/r2c_ncu_hdw.log 2>&1; echo "ncu rc=$?"
head -c 1500 gpurun_out/r2c_bench_c2.json; echo; head -c 700 gpurun_out/r2c_bench_c3.json; echo; head -c 400 gpurun_out/r2c_bench_c4.json; echo; head -c 400 gpurun_out/r2c_bench_c5.json
