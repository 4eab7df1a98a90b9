#!/usr/bin/env bash
# Builds librtw_b200.so (the C-ABI library) for sm_100a, in-tree, next to the sources.
#   -fmad=false : nothing is contracted unless written as fmaf()/fma() (FP contract)
#   -lineinfo   : ncu source view maps SASS to these files
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false
       -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v)
"$NVCC" "${FLAGS[@]}" -c rtw_kernels.cu -o rtw_kernels.o
"$NVCC" "${FLAGS[@]}" -c rtw_capi.cu -o rtw_capi.o
"$NVCC" "${FLAGS[@]}" -c rtw_wavefront.cu -o rtw_wavefront.o
"$NVCC" "${FLAGS[@]}" -c rtw_cta_wavefront.cu -o rtw_cta_wavefront.o
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a rtw_kernels.o rtw_capi.o rtw_wavefront.o rtw_cta_wavefront.o -o librtw_b200.so -lpthread -ldl
echo "built $(pwd)/librtw_b200.so"
