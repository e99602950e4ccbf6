#!/bin/bash
# tools/build_variant.sh NAME [nvcc -D flags]: a tuning variant of the product library (same C ABI) as gmrm_b200/variants/lib_NAME.so;
# GMRM_B200_LIB=<path> makes gmrm_b200/api.py (tests, bench) load it instead of the product build.
set -e
cd "$(dirname "$0")/../gmrm_b200/csrc"
NAME=$1; shift
O=_obj/var_$NAME; mkdir -p $O ../variants
for f in kernels predict engine; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c $f.cu -o $O/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$NAME.so $O/kernels.o $O/predict.o $O/engine.o -ldl
cuobjdump -res-usage $O/kernels.o 2>&1 | grep -A1 "step_kernelILi1" | grep REG
