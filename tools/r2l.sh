#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
{
tools/ab_env.sh C5 KTK_LIB=gpurun_variants/libktk_base.so
tools/ab_env.sh C5 KTK_X=intree
tools/ab_env.sh C5 KTK_LIB=gpurun_variants/libktk_base.so
tools/ab_env.sh C5 KTK_X=intree
BENCH_EXTRA="--row-order device" tools/ab_env.sh C5 KTK_LIB=gpurun_variants/libktk_base.so ROW=device
BENCH_EXTRA="--row-order device" tools/ab_env.sh C5 KTK_X=intree ROW=device
} 2>&1 | tee gpurun_out/r2y_split_tma_ab.log
