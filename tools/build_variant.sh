#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG=VALUE ...]  -> gpurun_variants/lib_NAME.so (experiment builds; not shipped)
N=$1; shift
mkdir -p gpurun_variants
nvcc -O3 -std=c++17 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo "$@" \
  -o gpurun_variants/lib_$N.so hardware-efficient-mua-compression_b200/csrc/mua_abi.cu
