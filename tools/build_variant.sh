#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG=VALUE ...]  -> gpurun_variants/libktk_NAME.so (git-ignored; travels with gpurun).
# A/B builds of the CUDA library for kernel-tuning experiments: run with KTK_LIB=gpurun_variants/libktk_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p gpurun_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -cudart shared \
  -Xlinker -rpath=/usr/local/cuda/lib64 "$@" -o gpurun_variants/libktk_$name.so kontiki_b200/csrc/ktk.cu
echo gpurun_variants/libktk_$name.so
