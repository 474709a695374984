#!/bin/bash
# Build liboqupy_b200.so for sm_100a in-tree (the .so travels to the GPU box).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared ${NVCC_EXTRA} \
  common.cu zgemm.cu svd.cu dyn.cu chain.cu batch.cu \
  -o ../liboqupy_b200.so
echo "built $(cd .. && pwd)/liboqupy_b200.so"
