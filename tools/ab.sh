#!/bin/bash
# tools/ab.sh [bench args]: bench the in-tree library and every variant under
# gpurun_variants/ (tools/build_variant.sh); one line each.
cd "$(dirname "$0")/.."
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py --quick --steps 20 --warmup 5 $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); k=l['roofline']['kernel_ms_per_step']
print('%-28s step %.4f ms  value %.3f G/s  cam %.4f  gyro %.4f  accel %.4f  frac %.3f  e2e %.1f M/s' % ('$label', l['ms_per_step'], l['value']/1e9, k.get('cam',0), k.get('gyro',0), k.get('accel',0), l['roofline']['frac'] or 0, l['e2e']['value']/1e6))"
}
BENCH_ARGS="$*"
run intree KTK_X=0
for f in gpurun_variants/libktk_*.so; do
  [ -e "$f" ] || continue
  run "$(basename $f .so)" KTK_LIB=$f
done
