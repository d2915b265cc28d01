#!/bin/bash
# round-2 GPU job Q (pair kernel): bench lines of C4 / C5 on one GPU, reference arm
mkdir -p gpurun_out
timeout 1200 python bench.py --config C4 --steps 3 --warmup 1 --no-cpu > gpurun_out/r2u_bench_c4.json 2> gpurun_out/r2u_bench_c4.err; echo "C4 rc=$?"; tail -c 300 gpurun_out/r2u_bench_c4.err
timeout 1200 python bench.py --config C5 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2u_bench_c5.json 2> gpurun_out/r2u_bench_c5.err; echo "C5 rc=$?"; tail -c 300 gpurun_out/r2u_bench_c5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2u_bench_reference.json 2> gpurun_out/r2u_bench_reference.err; echo "ref rc=$?"
for c in c4 c5; do python -c "
import json; d=json.loads(open('gpurun_out/r2u_bench_$c.json').read().strip().splitlines()[-1]); print('$c', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks'])"; done
