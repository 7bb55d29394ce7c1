#!/bin/bash
# tools/final_run.sh TAG: end-of-round evidence on one B200 -- GPU tests, smoke, every bench line, the reference arm, launch list, ncu --set full.
TAG=${1:-r2z}
cd "$(dirname "$0")/.."
O=gpurun_out/final_$TAG
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee $O/gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
b() { name=$1; shift; python bench.py "$@" 2>$O/$name.err | tail -1 > $O/$name.json; python - <<PY
import json
try:
    l = json.loads(open("$O/$name.json").read()); r = l.get("roofline") or {}
    print("%-12s ms/step %.4f  value %.4g  e2e %.4g  frac %s  kernel %s  parity %s" % ("$name", l["ms_per_step"], l["value"], (l.get("e2e") or {}).get("value", 0), r.get("frac"), r.get("kernel"), (l.get("parity") or {}).get("pass")))
except Exception as e:
    print("$name failed:", e)
PY
}
b h1
b h1_dev --row-order device --no-cpu-baseline
b c1 --workload C1
b c2 --workload C2
b c3 --workload C3 --no-cpu-baseline
b c4 --workload C4 --no-cpu-baseline
b c5 --workload C5
b c3_newton --workload C3 --camera-method newton --no-cpu-baseline --steps 20
b c3_lifting --workload C3 --camera-method lifting --no-cpu-baseline --steps 50
b c3_atan --workload C3 --camera-model atan --no-cpu-baseline
b strong_c4_n1 --workload C4 --scaling strong --no-cpu-baseline
b strong_c5_n1 --workload C5 --scaling strong --no-cpu-baseline
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
KTK_FUSE_SHORT=0 ncu --set full --clock-control none --import-source on -k regex:"k_static_rs|k_landmark_ref|k_imu|k_pair_prepass" -s 12 -c 5 -f -o $O/full python bench.py --quick --no-cpu-baseline --steps 2 --warmup 3 > $O/ncu_full.log 2>&1
ls -la $O | tail -30
