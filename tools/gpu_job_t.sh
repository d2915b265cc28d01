#!/bin/bash
mkdir -p gpurun_out
LDW_DBG_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu --no-extra --no-post > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo rc=$?
grep "ldw timing" gpurun_out/t_bench.err | tail -60
