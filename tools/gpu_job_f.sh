#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2f_gpu_tests.log 2>&1; tail -6 gpurun_out/r2f_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err; echo "C2 rc=$?"; tail -c 300 gpurun_out/r2f_bench_c2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['detail']['wall_to_links_s']); print(json.dumps(d['detail']['extra'])[:3000])"
