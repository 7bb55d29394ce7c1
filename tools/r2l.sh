#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for wl in C4 C2 C1 H1; do for f in 0 1 0 1; do tools/ab_env.sh $wl KTK_FUSE_SHORT=$f; done; done
} 2>&1 | tee gpurun_out/r2x_fuse_rule_ab.log
