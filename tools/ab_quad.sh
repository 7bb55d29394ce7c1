#!/bin/bash
# tools/ab_quad.sh: parity of the four-lanes-per-row camera kernel on the GPU, then A/B of it (env switch KTK_CAM_QUAD on the in-tree
# build; occupancy variants from tools/build_variant.sh under gpurun_variants/) against the one-thread-per-row kernel.
cd "$(dirname "$0")/.."
KTK_CAM_QUAD=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py --no-cpu-baseline --steps 20 --warmup 5 $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); k=l['roofline']['kernel_ms_per_step']
print('%-28s step %.4f ms  value %.3f G/s  cam %.4f  gyro %.4f  accel %.4f  frac %.3f' % ('$label', l['ms_per_step'], l['value']/1e9, k.get('cam',0), k.get('gyro',0), k.get('accel',0), l['roofline']['frac'] or 0))"
}
for BENCH_ARGS in "" "--row-order device"; do
  echo "== bench.py $BENCH_ARGS"
  run row KTK_CAM_QUAD=0
  run quad12 KTK_CAM_QUAD=1
  for f in gpurun_variants/libktk_*.so; do
    [ -e "$f" ] || continue
    run "$(basename $f .so)" KTK_LIB=$f
  done
done
