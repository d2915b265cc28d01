#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2e_gpu_tests.log 2>&1; tail -6 gpurun_out/r2e_gpu_tests.log
timeout 600 python bench.py --config C3 --steps 5 --warmup 2 > gpurun_out/r2e_bench_c3.json 2> gpurun_out/r2e_bench_c3.err; echo "C3 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_c3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'])"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_c2.json 2> gpurun_out/r2e_bench_c2.err; echo "C2 rc=$?"; tail -c 300 gpurun_out/r2e_bench_c2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['detail']['wall_to_links_s'], d['detail']['extra'])"
