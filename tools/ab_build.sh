#!/bin/bash
# tools/ab_build.sh NAME "-DFLAG=.. ..." : builds rnacode_b200/lib/ab/libRNAcode_cuda_NAME.so with extra nvcc flags
# (A/B experiments; select with RNACODE_CUDA_LIB=... when running bench.py)
set -e
cd "$(dirname "$0")/.."
mkdir -p rnacode_b200/lib/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -fmad=false -prec-div=true -prec-sqrt=true $2 \
  -o rnacode_b200/lib/ab/libRNAcode_cuda_$1.so rnacode_b200/csrc/rnacode_cuda.cu
