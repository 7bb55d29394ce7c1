#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
{
for wl in C2 H1 C1; do
tools/ab_env.sh $wl KTK_LIB=gpurun_variants/libktk_imuscatter.so
tools/ab_env.sh $wl KTK_X=intree
tools/ab_env.sh $wl KTK_LIB=gpurun_variants/libktk_imuscatter.so
tools/ab_env.sh $wl KTK_X=intree
done
} 2>&1 | tee gpurun_out/r2y_imu_tma_ab.log
