#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_gpu_tests.log 2>&1; tail -6 gpurun_out/r2d_gpu_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'column_count_kernel|extract_codes_kernel|hdw_gemm_kernel|hdw_pack_kernel|snp_allele_stats_kernel|site_filter_kernel' -c 40 --csv --log-file gpurun_out/r2d_stage_launches.csv python tools/bench_stages.py > gpurun_out/r2d_stage_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/bench_stages.py > gpurun_out/r2d_stage_timings.jsonl 2> gpurun_out/r2d_stage.err; cat gpurun_out/r2d_stage_timings.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err; echo "C2 rc=$?"; tail -c 300 gpurun_out/r2d_bench_c2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['breakdown_ms'])"
