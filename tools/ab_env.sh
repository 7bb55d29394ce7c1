#!/bin/bash
# tools/ab_env.sh WORKLOAD VAR=VAL ... : one quick bench line (step time, per-kernel times) under the given environment switches
cd "$(dirname "$0")/.."
wl=$1; shift
env "$@" python bench.py --quick --steps 50 --warmup 5 --workload $wl $BENCH_EXTRA 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); k=l['roofline']['kernel_ms_per_step']
print('%-6s %-22s step %.4f ms (median %.4f)  value %.3f G/s  cam %.4f  gyro %.4f  accel %.4f  frac %.3f' % ('$wl', '$*', l['ms_per_step'], l['timing']['ms_per_step_median'], l['value']/1e9, k.get('cam',0), k.get('gyro',0), k.get('accel',0), l['roofline']['frac'] or 0))"
