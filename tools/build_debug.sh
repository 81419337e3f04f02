#!/bin/bash
# Builds pesr_b200/libpesr_b200_debug.so: the product sources + the bring-up hooks (include/pesr_b200_debug.h).
# The tools that need the hooks load it with PESR_B200_LIB=pesr_b200/libpesr_b200_debug.so (see pesr_b200/_debug.py).
set -e
cd "$(dirname "$0")/.."
${NVCC:-/usr/local/cuda/bin/nvcc} -shared -Xcompiler -fPIC -std=c++17 -O3 -lineinfo -DPESR_DEBUG_HOOKS \
  -gencode arch=compute_100a,code=sm_100a -o pesr_b200/libpesr_b200_debug.so pesr_b200/csrc/*.cu tools/csrc_debug/*.cu
echo built pesr_b200/libpesr_b200_debug.so
