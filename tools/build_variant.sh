#!/bin/bash
# Builds a kernel-experiment variant of the library: tools/build_variant.sh NAME [-DFLAG=VALUE ...]
# -> ros2-recursive-patchwork-implementation_b200/_variants/NAME.so  (timed by tools/gpu_variants.py)
set -e
cd "$(dirname "$0")/../ros2-recursive-patchwork-implementation_b200"
mkdir -p _variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -I../include -Icsrc "$@" \
     -shared -o _variants/$name.so csrc/rpw_kernels.cu csrc/rpw_capi.cu 2>&1 | grep -v "warning #177\|declared but never\|constexpr int\|\^\|^$\|Remark" || true
ls -la _variants/$name.so
