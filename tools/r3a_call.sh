python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 2 3; do KTK_NEWTON_FAST=$f python bench.py --workload C3 --camera-method newton --no-cpu-baseline --steps 20 2>/dev/null | tail -1 > gpurun_out/r3a_newton_fast$f.json; python - <<PY
import json
l=json.loads(open("gpurun_out/r3a_newton_fast$f.json").read())
print("KTK_NEWTON_FAST=$f ms/step %.4f value %.4g launches %s parity %s" % (l["ms_per_step"], l["value"], l.get("gpu_launches"), l.get("parity",{}).get("pass")))
PY
done
