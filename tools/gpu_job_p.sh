#!/bin/bash
# round-2 GPU job P (pair kernel): launch list + full capture of the scan kernel for profiles/, sanitizer on the smoke
mkdir -p gpurun_out
T=${1:-r2p}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/${T}_launches_run.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mi_scan_kernel -s 30 -c 1 -f -o gpurun_out/${T}_mi_scan python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/${T}_full_run.log 2>&1; echo "full rc=$?"
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${T}_memcheck.log
