#!/bin/bash
cd "$(dirname "$0")/.."
N=$1
mkdir -p gpurun_out
for wl in C4 C5; do
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $wl --scaling strong --no-cpu-baseline 2>gpurun_out/r2z_strong_${wl}_n$N.err | tail -1 > gpurun_out/r2z_strong_${wl}_n$N.json
python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2z_strong_${wl}_n$N.json").read()); s = l["strong"]
    print("$wl N=$N ms/step %.4f value %.4g e2e %.4g numa %s graphs %s vs_unsharded_ok %s parity %s" % (l["ms_per_step"], l["value"], l["e2e"]["value"], l["e2e"].get("numa"), s.get("cuda_graphs_between_collectives"), (s.get("vs_unsharded") or {}).get("ok"), (l.get("parity") or {}).get("pass")), s["ms"], s["allreduce_ms"])
except Exception as e:
    print("$wl failed:", e); print(open("gpurun_out/r2z_strong_${wl}_n$N.err").read()[-1500:])
PY
done
