# round-3 session call B: the closed-form Newton path on the GPU (tests, bench, launch list) + A/B of the sequence points (KB_SEQ) in the accelerometer / lifting rows
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "newton or span or lifting" 2>&1 | tail -3
timeout 60 python bench.py --workload C3 --camera-method newton --no-cpu-baseline --steps 20 2>/dev/null | tail -1 > gpurun_out/r3b_newton_fast4.json; python - <<PY
import json
l=json.loads(open("gpurun_out/r3b_newton_fast4.json").read())
print("KTK_NEWTON_FAST=4 ms/step %.4f value %.4g launches %s parity %s" % (l["ms_per_step"], l["value"], l.get("gpu_launches"), l.get("parity",{}).get("pass")))
PY
timeout 60 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none -c 40 --csv --log-file gpurun_out/r3b_newton_launches.csv python bench.py --workload C3 --camera-method newton --no-cpu-baseline --quick --steps 2 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r3b_newton_launches.csv")) if len(r)>14 and r[12]=="gpu__time_duration.sum"]
for r in rows[-10:]: print(r[4][:70], r[14])
PY
(echo "== H1"; timeout 60 bash tools/ab.sh; echo "== C2"; timeout 60 bash tools/ab.sh --workload C2; echo "== C3 lifting"; timeout 60 bash tools/ab.sh --workload C3 --camera-method lifting) 2>&1 | tee gpurun_out/r3b_seq_ab.log
