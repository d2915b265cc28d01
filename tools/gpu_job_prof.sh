#!/bin/bash
# full ncu capture of one mi_scan_kernel launch (off-diagonal 10k x 10k block of the 616 x 30000 A/B workload)
mkdir -p gpurun_out
python tools/kernel_ab.py --nsnp 30000 >/dev/null 2>&1   # writes /tmp/kernel_ab_data.npz
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mi_scan_kernel -s 26 -c 2 -f -o gpurun_out/${1:-prof}_mi_scan python tools/kernel_ab.py --worker base --data /tmp/kernel_ab_data.npz --steps 1 > gpurun_out/${1:-prof}_run.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/${1:-prof}_run.log | cut -c1-300
