#!/bin/bash
# full single-GPU validation of the current library: determinism stress, the whole -m gpu suite, bench C2 + reference arm
mkdir -p gpurun_out
T=${1:-r2p}
REPS=30 timeout 900 python tools/stress_determinism.py 30 2>&1 | tail -5 | tee gpurun_out/${T}_stress.log
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench_c2.err
python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['detail'].get('wall_to_links_s'), d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
