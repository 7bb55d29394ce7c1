#!/bin/bash
# tools/scale_run.sh N TAG: the bench under torchrun on N GPUs of one box -- weak scaling (H1, one full problem per rank, no data-path collective)
# and strong scaling (C4, C5: ONE problem, rows sharded, Gauss-Newton buffers all-reduced over NCCL); one JSON line each under gpurun_out/.
cd "$(dirname "$0")/.."
N=$1; TAG=${2:-r2}
mkdir -p gpurun_out
run() { # name, bench args...
  name=$1; shift
  if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@" 2>gpurun_out/${TAG}_${name}_n$N.err | tail -1 > gpurun_out/${TAG}_${name}_n$N.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>gpurun_out/${TAG}_${name}_n$N.err | tail -1 > gpurun_out/${TAG}_${name}_n$N.json; fi
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/${TAG}_${name}_n$N.json").read())
    s = l.get("strong") or {}
    print("${name} N=$N", l["scaling"], "ms/step %.4f" % l["ms_per_step"], "value %.4g" % l["value"], "e2e %.4g" % (l.get("e2e") or {}).get("value", 0), "allreduce_ms", s.get("allreduce_ms"), "ms", s.get("ms"), "parity", l.get("parity"))
except Exception as e:
    print("${name} N=$N failed:", e); print(open("gpurun_out/${TAG}_${name}_n$N.err").read()[-1500:])
PY
}
run weak_H1 --no-cpu-baseline
run strong_C4 --workload C4 --scaling strong --no-cpu-baseline
run strong_C5 --workload C5 --scaling strong --no-cpu-baseline
