# round-3 session call C (last GPU seconds of the round): full GPU suite + smoke on the shipped defaults, LiftingRs rows built in place (A/B),
# NewtonRs closed form with the early exit of k_newton_rs, H1 headline
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r3c_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r3c_smoke.txt
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[2]).read()); r = l.get("roofline") or {}
    print("%-22s ms/step %.4f  value %.4g  frac %s  kernels %s  parity %s" % (sys.argv[1], l["ms_per_step"], l["value"], r.get("frac"), r.get("kernel_ms_per_step"), (l.get("parity") or {}).get("pass")))
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
}
for m in 0 1; do KTK_LIFT_TMA=$m timeout 40 python bench.py --workload C3 --camera-method lifting --quick --no-cpu-baseline --steps 50 2>/dev/null | tail -1 > gpurun_out/r3c_lifting_tma$m.json; line "lifting KTK_LIFT_TMA=$m" gpurun_out/r3c_lifting_tma$m.json; done
timeout 40 python bench.py --workload C3 --camera-method newton --quick --no-cpu-baseline --steps 50 2>/dev/null | tail -1 > gpurun_out/r3c_newton.json; line "newton (default 4)" gpurun_out/r3c_newton.json
timeout 60 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r3c_h1.json; line "H1" gpurun_out/r3c_h1.json
