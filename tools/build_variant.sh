#!/bin/bash
# Developer tool: build a variant of libldwgpu.so whose scan kernel is compiled with extra -D flags, for same-box A/B
# timing (tools/kernel_ab.py).  usage: tools/build_variant.sh NAME [-DFLAG=1 ...]   ->  ldweaver_b200/variants/libldwgpu_NAME.so
set -e
name=$1; shift
cd "$(dirname "$0")/../ldweaver_b200/csrc"
make -s >/dev/null
mkdir -p build ../variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-Wall -Xptxas -v "$@" -c -o build/mi_scan_$name.o mi_scan.cu 2> build/mi_scan_$name.ptxas.log || { cat build/mi_scan_$name.ptxas.log; exit 1; }
grep -A2 "mi_scan_kernelILb0" build/mi_scan_$name.ptxas.log | grep -E "spill|Used" || true
objs=$(ls build/*.o | grep -v "mi_scan" | tr '\n' ' ')
$NVCC $ARCH -shared -o ../variants/libldwgpu_$name.so $objs build/mi_scan_$name.o -lz -ldl
echo "built ldweaver_b200/variants/libldwgpu_$name.so"
