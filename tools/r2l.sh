#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
{
tools/ab_env.sh C5 KTK_X=intree
} 2>&1 | tee gpurun_out/r2y_c5_imu_tma.log
