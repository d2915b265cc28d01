#!/bin/bash
# round-2 GPU job K: launch list + full capture of the scan kernel for profiles/, bench line, wall tool
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2k_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/r2k_launches_run.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mi_scan_kernel -s 30 -c 1 -o gpurun_out/r2k_mi_scan python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/r2k_full_run.log 2>&1; echo "full rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2k_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err; echo "ref rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2k_bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['detail']['wall_to_links_s'], d['roofline']['frac'], d['roofline']['mufu']['frac']); print(json.dumps(d['detail']['extra'])[:2500])
r=json.loads(open('gpurun_out/r2k_bench_reference.json').read().strip().splitlines()[-1]); print('reference', r['value'], r['cpu_baseline']['cores'])"
