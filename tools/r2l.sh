#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
tools/ab_env.sh H1 KTK_LIB=gpurun_variants/libktk_l1win0.so
tools/ab_env.sh H1 KTK_X=intree
tools/ab_env.sh H1 KTK_LIB=gpurun_variants/libktk_l1win3.so
tools/ab_env.sh H1 KTK_LIB=gpurun_variants/libktk_l1win0.so
tools/ab_env.sh H1 KTK_X=intree
tools/ab_env.sh H1 KTK_LIB=gpurun_variants/libktk_l1win3.so
BENCH_EXTRA="--row-order device" tools/ab_env.sh H1 KTK_X=intree ROW=device
} 2>&1 | tee gpurun_out/r2w_l1win_ab2.log
