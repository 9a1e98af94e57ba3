#!/bin/bash
# Builds libgmr_b200.so for sm_100a in-tree (the .so travels to the GPU box with the repo).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
  -Xptxas -v -Xcompiler -fPIC -shared -o libgmr_b200.so gmr_kernels.cu "$@"
