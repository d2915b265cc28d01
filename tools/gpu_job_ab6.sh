#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kernel_ab.py pair "$@" pair "$@" --nsnp 30000 2>&1 | cut -c1-100 | tee gpurun_out/ab6.log
