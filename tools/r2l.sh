#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
{
for wl in H1 C3; do
  tools/ab_env.sh $wl KTK_LIB=gpurun_variants/libktk_base.so
  tools/ab_env.sh $wl KTK_X=intree
done
BENCH_EXTRA="--row-order device" tools/ab_env.sh H1 KTK_LIB=gpurun_variants/libktk_base.so ROW=device
BENCH_EXTRA="--row-order device" tools/ab_env.sh H1 KTK_X=intree ROW=device
} 2>&1 | tee gpurun_out/r2s_park_ab.log
