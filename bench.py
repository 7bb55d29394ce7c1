#!/usr/bin/env python
"""bench.py -- measurements/s of one full residual + Jacobian evaluation (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload H1]

A "step" is one evaluation of every measurement of the workload at a fixed parameter point.
  value        : device-resident throughput (inputs, knots and outputs in HBM; K steps between two CUDA events on the
                 launching stream, barrier + synchronize on both sides, max over ranks)
  e2e          : the same through ktk_evaluate with HOST buffers (H2D of the parameter point, D2H of every residual and
                 Jacobian row inside the timed region)
  roofline     : dominant kernel (static-RS rows) timed with CUDA events inside the library, against MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (restated reference, Ceres-style stride-4 autodiff, all host threads) on a bounded sample
With N > 1 (torchrun) every rank evaluates its own full-size shard (measurements are independent: no data-path
collective, weak scaling).  --impl reference times the CPU oracle only (the reference itself cannot be built here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NTHREADS = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)      # the CPU arm uses every host core explicitly (torchrun exports OMP_NUM_THREADS=1)
METRIC = "measurements/s (residual+Jacobian eval)"
UNIT = "measurements/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="H1")
    ap.add_argument("--cpu-sample", type=float, default=None, help="fraction of the workload the CPU baseline evaluates per step (default: sized for ~12 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): every rank evaluates its own full-size set of measurements, no data-path collective.  "
                         "strong (BASELINE.json configs[3]/[4]): ONE problem, rows sharded over the ranks by kontiki_b200/sharding.py, knots and inverse depths "
                         "replicated; a step = evaluation + J^T r + one (J^T J) v with the NCCL all-reduce of the parameter-sized vector")
    ap.add_argument("--quick", action="store_true", help="kernel-tuning runs (tools/ab.sh): device-resident leg and per-kernel times only; no e2e leg, no CPU baseline")
    ap.add_argument("--camera-method", default="static", choices=["static", "newton", "lifting"],
                    help="StaticRsCameraMeasurement (the BASELINE.json workloads) or NewtonRsCameraMeasurement rows (SURVEY.md 8f-3) for the camera group")
    ap.add_argument("--camera-model", default="pinhole", choices=["pinhole", "atan"], help="PinholeCamera (BASELINE.json) or AtanCamera")
    ap.add_argument("--row-order", default="caller", choices=["caller", "device"],
                    help="row order of the device-resident leg: the caller's insertion order (default, what the drop-in contract documents) or "
                         "KTK_EVAL_DEVICE_ORDER (rows sorted by first knot, one TMA bulk store per warp tile)")
    return ap.parse_args()


ATAN = dict(wc=(0.0029110778971412417, 0.0004189670467132041), gamma=0.8894355177968156)      # python/tests/fixtures/camera_fixtures.py:15-16


def workload_config(name, cfg, row_order="caller", method="static", model="pinhole", strong=False, world=1):
    from kontiki_b200 import synthetic as syn
    traj = "SplitTrajectory (UniformR3 + UniformSO3)" if cfg.get("split") else "UniformSE3SplineTrajectory"
    return {"workload": f"{name}: {traj} {len(cfg['knots'])} knots dt={cfg['dt']}, "
                        f"{len(cfg['gyro']['t']) if cfg['gyro'] else 0} gyro + {len(cfg['accel']['t']) if cfg['accel'] else 0} accel (BasicImu) + "
                        f"{len(cfg['cam']['lm_idx']) if cfg['cam'] else 0} {dict(newton='NewtonRsCamera', lifting='LiftingRsCamera').get(method, 'StaticRsCamera')} "
                        f"({'Atan' if model == 'atan' else 'Pinhole'}, {len(cfg['cam']['rho']) if cfg['cam'] else 0} landmarks)",
            ("measurements_per_step_total" if strong else "measurements_per_step_per_gpu"): syn.num_measurements(cfg),
            ("algorithmic_bytes_per_step_total" if strong else "algorithmic_bytes_per_step_per_gpu"): syn.algorithmic_bytes(cfg),
            "jacobian": "ambient (7 per SE3 knot; 3 + 4 per split knot), Huber corrector applied to camera rows",
            "l2": "per-step working set (outputs + records) >> 126 MB L2; no flush",
            "sharding": ("ONE problem: rows sharded over %d rank(s) (IMU rows by time range, camera rows by landmark), knots / inverse depths replicated; step = evaluation + "
                         "J^T r + (J^T J) v, one NCCL all-reduce of the parameter vector per product" % world) if strong else
                        "measurements sharded across ranks, knots replicated, no data-path collective",
            "row_order": row_order}


# ---- CPU oracle leg (cpu_baseline and --impl reference) ------------------------------------------------------------
def oracle_sample(cfg, frac, seed=0):
    """A bounded random sample of the workload with the same mix of measurement types."""
    rng = np.random.default_rng(seed)
    out = {}
    for k in ("gyro", "accel"):
        if cfg[k]:
            n = len(cfg[k]["t"])
            sel = rng.permutation(n)[:max(1, int(n * frac))]
            out[k] = {a: v[sel] for a, v in cfg[k].items()}
    if cfg["cam"]:
        c = cfg["cam"]
        n = len(c["lm_idx"])
        sel = rng.permutation(n)[:max(1, int(n * frac))]
        out["cam"] = dict(c, **{a: c[a][sel] for a in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "weight", "huber_c")})
    return out


CAMERA = {"method": "static", "model": "pinhole"}      # set from the command line in main()


def oracle_step(cfg, sample):
    """One residual+Jacobian evaluation of the sample with the CPU oracle; returns (rows, seconds inside Evaluate)."""
    from oracle import kto
    if cfg.get("split"):
        traj = kto.Traj(kto.SPLIT, cfg["dt"], cfg["t0"], cfg["r3"], cfg["dt"], cfg["t0"], cfg["so3"])
    else:
        traj = kto.Traj(kto.SE3, cfg["dt"], cfg["t0"], cfg["knots"])
    rows, secs = 0, 0.0
    imu = kto.Sensor()
    for which, k in ((0, "gyro"), (1, "accel")):
        if k in sample:
            m = sample[k]
            res = kto.imu_residuals(traj, imu, which, m["t"], m["y"], m["weight"], jac_mode=1, nthreads=NTHREADS)
            rows += len(m["t"]); secs += res["eval_seconds"]
    if "cam" in sample:
        c = sample["cam"]
        ocam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"], method="static" if CAMERA["method"] == "lifting" else CAMERA["method"],
                          **(ATAN if CAMERA["model"] == "atan" else {}))
        if CAMERA["method"] == "lifting":
            res = kto.lifting_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], weight=c["weight"], jac_mode=1,
                                           cap=24, nthreads=NTHREADS)
        else:
            res = kto.static_rs_residuals(traj, ocam, c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["rho"], c["weight"], jac_mode=1, cap=24,
                                          nthreads=NTHREADS)
        rows += len(c["lm_idx"]); secs += res["eval_seconds"]
    return rows, secs


def cpu_baseline(cfg, frac, steps=1, warmup=1, budget_s=12.0):
    from oracle import kto
    if frac is None:      # pilot on 0.5 % of the workload, then size the sample for ~budget_s seconds of CPU work in total
        r, s = oracle_step(cfg, oracle_sample(cfg, 0.005, seed=1))
        n_total = sum(len(cfg[k]["t"]) for k in ("gyro", "accel") if cfg[k]) + (len(cfg["cam"]["lm_idx"]) if cfg["cam"] else 0)
        frac = float(min(1.0, max(0.005, budget_s / max(steps + warmup, 1) * (r / s) / n_total)))
    sample = oracle_sample(cfg, frac)
    for _ in range(warmup):
        oracle_step(cfg, sample)
    rows, secs = 0, 0.0
    for _ in range(steps):
        r, s = oracle_step(cfg, sample)
        rows += r; secs += s
    base = {"value": rows / secs, "unit": UNIT, "cores": NTHREADS, "kind": "port",
            "sample": f"{rows // steps} rows per step ({frac:.3g} of the workload, same type mix), {steps} step(s), "
                      "time inside the per-block Evaluate loop only (problem construction excluded), OpenMP over blocks"}
    opt = optimised_cpu(cfg, sample)
    if opt:
        base["optimised"] = opt
        # the driver keeps value / unit / cores / kind / sample: the optimised-CPU figure rides along in `sample` as well
        base["sample"] += (f"; OPTIMISED CPU variant on the same sample (analytic Jacobians, hoisted knot-pair log, oracle/analytic_cpu.cpp, {opt['cores']} threads): "
                           f"{opt['value']:.4g} {UNIT}")
    return base, rows, secs


def optimised_cpu(cfg, sample, repeats=3):
    """next to the restated reference (autodiff multipass per block, above) report an OPTIMISED CPU variant -- analytic
    Jacobians, knot-pair log hoisted -- so that the GPU / CPU ratio is not inflated by autodiff overhead alone.  oracle/analytic_cpu.cpp:
    the product's own closed-form mathematics compiled for the host, OpenMP over rows, same sample.  SE3 + static pinhole / atan rows only."""
    from oracle import kto
    if cfg.get("split") or CAMERA["method"] != "static":
        return None
    cam = None
    if "cam" in sample:
        c = sample["cam"]
        cam = {k: c[k] for k in ("K", "readout", "rows", "obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "rho", "weight", "huber_c")}
        if CAMERA["model"] == "atan":
            cam.update(ATAN)
    best, rows = None, 0
    for _ in range(repeats + 1):                       # first pass warms the caches / thread pool
        res = kto.analytic_se3_evaluate(cfg["dt"], cfg["t0"], cfg["knots"], gyro=sample.get("gyro"), accel=sample.get("accel"), cam=cam, nthreads=NTHREADS)
        rows = sum(len(sample[k]["t"]) for k in ("gyro", "accel") if k in sample) + (len(sample["cam"]["lm_idx"]) if "cam" in sample else 0)
        best = res["seconds"] if best is None else min(best, res["seconds"])
    return {"value": rows / best, "unit": UNIT, "cores": NTHREADS, "kind": "port-analytic",
            "what": "oracle/analytic_cpu.cpp: the product's closed-form row mathematics (analytic SE(3) Jacobians, knot-pair log hoisted into one prepass) "
                    "compiled for the host, OpenMP over rows; landmark side per measurement; best of %d passes over the same sample" % repeats}



# ---- parity gate (SURVEY.md 8d "Parity gate run with every benchmark"): the oracle as the CHECKER, after the timed region ------
def parity_gate(cfg, p, groups, keep, rho, flags, device_order, n_sample=2048, seed=7):
    """Samples n_sample rows per group from the DEVICE-RESIDENT outputs of the benchmark's last step and compares them with the CPU oracle
    (oracle/gate.py): indices bit-exact, residuals / Jacobians relative to the row's largest entry (camera residuals: absolute, in pixels)."""
    import torch
    from oracle import gate
    from kontiki_b200 import _lib
    out = {"rows_checked": 0, "idx_exact": True, "max_rel_r": 0.0, "max_rel_J": 0.0, "max_abs_r_cam_px": 0.0, "tol": gate.TOL, "rows_per_group": n_sample,
           "checker": "oracle/ (restated reference, dual-number autodiff; pinned by the 60-digit transcription tests/mp_reference.py)"}
    rng = np.random.default_rng(seed)
    for name, g in groups.items():
        if name == "cam" and CAMERA["method"] != "static" and cfg.get("split"):
            out["camera"] = "not gated here (NewtonRs / LiftingRs rows on a split trajectory: tests/test_gpu_parity.py)"
            continue
        r_t, J_t, idx_t = keep[g]
        n = p.group_size(g)
        sel = np.sort(rng.permutation(n)[:min(n, n_sample)])
        pos = sel
        if device_order:
            inv = np.empty(n, np.int64)
            inv[p.get_row_order(g)] = np.arange(n)
            pos = inv[sel]
        tpos = torch.from_numpy(np.ascontiguousarray(pos)).to(r_t.device)
        try:
            res = gate.check_rows(cfg, name, sel, r_t.index_select(0, tpos).cpu().numpy(), J_t.index_select(0, tpos).cpu().numpy(),
                                  [t.index_select(0, tpos).cpu().numpy() for t in idx_t], rho, robust=bool(flags & _lib.EVAL_ROBUST),
                                  atan=ATAN if CAMERA["model"] == "atan" else None, nthreads=NTHREADS,
                                  **(dict(method=CAMERA["method"]) if name == "cam" else {}))      # NewtonRs / LiftingRs rows: span layout, the oracle differentiates through the iteration
        except Exception as e:      # the checker failing is a failed gate, not a lost benchmark line
            if name != "cam" or CAMERA["method"] == "static":
                raise
            out["camera"] = f"gate raised {type(e).__name__}: {e}"
            out["idx_exact"] = False
            continue
        out["idx_exact"] &= res["idx_exact"]
        for k_out, k_in in (("max_rel_r", "rel_r"), ("max_rel_J", "rel_J"), ("max_abs_r_cam_px", "abs_r_cam_px")):
            out[k_out] = max(out[k_out], res[k_in])
        out["rows_checked"] += len(sel)
    out["pass"] = bool(out["idx_exact"] and out["max_rel_r"] <= gate.TOL and out["max_rel_J"] <= gate.TOL and out["max_abs_r_cam_px"] <= gate.CAM_R_TOL)
    return out


# ---- clocks --------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop, self.th = index, [], False, None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity), so that the pinned host buffers of the end-to-end leg are
    allocated on the GPU's own NUMA node: with 8 ranks each copying 538 MB per step, buffers on the far socket put every byte on the
    inter-socket link (round 1: 1.76x end to end on 8 GPUs).  Returns a short description for the JSON line; never fails the run."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use or use == allowed:
            return {"bound": False, "cpus_allowed": len(allowed)}
        os.sched_setaffinity(0, use)
        return {"bound": True, "cpus": len(use), "cpus_allowed": len(allowed)}
    except Exception as e:      # no NVML, no affinity syscall, container cpuset: run unbound
        return {"bound": False, "why": str(e)[:80]}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from kontiki_b200 import synthetic as syn
    CAMERA.update(method=a.camera_method, model=a.camera_model)

    if a.impl == "reference":
        if rank != 0:
            return
        cfg = syn.make_config(a.workload)
        base, rows, secs = cpu_baseline(cfg, a.cpu_sample, steps=a.steps, warmup=1 if a.warmup > 0 else 0, budget_s=90.0)
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * secs / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(a.workload, cfg, "caller", a.camera_method, a.camera_model), "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "restated reference on host cores: the reference needs Ceres 1.x + Sophus + Eigen, none installable here (oracle/README.md)"}
        print(json.dumps(line))
        return

    # keep stdout clean for the ONE JSON line: libraries (NCCL's version banner) write to fd 1 during the run
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from kontiki_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (kontiki_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    # before any pinned allocation: the end-to-end leg is a host-memory / PCIe measurement.  One rank keeps every core (its CPU baseline uses them).
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else {"bound": False, "why": "single rank"}
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = syn.make_config(a.workload)
    strong = a.scaling == "strong"
    cfg_full = cfg
    if strong:
        from kontiki_b200 import sharding
        cfg = sharding.shard_config(cfg_full, rank, world)      # this rank's rows of the ONE problem
    if world > 1 and not strong:      # every rank owns a different, equally sized shard of measurements (independent seeds), knots replicated
        n_knots = len(cfg["knots"])
        if cfg["gyro"]:
            cfg["gyro"] = syn.make_imu(len(cfg["gyro"]["t"]), n_knots, cfg["dt"], seed=100 + rank)
        if cfg["accel"]:
            cfg["accel"] = syn.make_imu(len(cfg["accel"]["t"]), n_knots, cfg["dt"], seed=200 + rank, accel=True)
        if cfg["cam"]:
            cfg["cam"] = syn.make_static_rs(cfg["knots"], cfg["dt"], len(cfg["cam"]["rho"]), seed=300 + rank)
    n_meas = syn.num_measurements(cfg)
    alg_bytes = syn.algorithmic_bytes(cfg)

    p = _lib.Problem(local_rank)
    stream = torch.cuda.Stream()             # a real (non-default) stream: the library replays each step as one CUDA graph on it
    torch.cuda.set_stream(stream)
    p.set_stream(stream.cuda_stream)
    if cfg.get("split"):
        p.set_split_spline(cfg["dt"], cfg["t0"], len(cfg["r3"]), cfg["dt"], cfg["t0"], len(cfg["so3"]))
        knots_flat = np.concatenate([cfg["r3"].reshape(-1), cfg["so3"].reshape(-1)])
    else:
        p.set_se3_spline(cfg["dt"], cfg["t0"], len(cfg["knots"]))
        knots_flat = cfg["knots"].reshape(-1)
    imu = _lib.make_sensor()
    groups = {}
    if cfg["gyro"]:
        groups["gyro"] = p.add_gyroscope(imu, cfg["gyro"]["t"], cfg["gyro"]["y"], cfg["gyro"]["weight"])
    if cfg["accel"]:
        groups["accel"] = p.add_accelerometer(imu, cfg["accel"]["t"], cfg["accel"]["y"], cfg["accel"]["weight"])
    rho = None
    if cfg["cam"]:
        c = cfg["cam"]
        rho = c["rho"]
        add = {"newton": p.add_newton_rs, "lifting": p.add_lifting_rs}.get(a.camera_method, p.add_static_rs)
        groups["cam"] = add(_lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"], **(ATAN if a.camera_model == "atan" else {})), c["obs_uv"], c["obs_t0"],
                            c["ref_uv"], c["ref_t0"], c["lm_idx"], c["weight"], c["huber_c"])
    flags = _lib.EVAL_RESIDUALS | _lib.EVAL_JACOBIANS | _lib.EVAL_ROBUST

    # ---- device-resident leg -------------------------------------------------------------------------------------
    dev = torch.device("cuda", local_rank)
    d_knots = torch.from_numpy(knots_flat).to(dev)
    d_rho = torch.from_numpy(rho).to(dev) if rho is not None else None
    d_outs, keep = [], []
    for g in range(p.num_groups):
        n, cam = p.group_size(g), p.group_kind(g) in (_lib.STATIC_RS, _lib.NEWTON_RS, _lib.LIFTING_RS)
        r = torch.empty((n, 2 if p.group_kind(g) in (_lib.STATIC_RS, _lib.NEWTON_RS) else 3), dtype=torch.float64, device=dev)
        J = torch.empty((n, p.group_row_size(g)), dtype=torch.float64, device=dev)
        idx = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4)]
        keep.append((r, J, idx))
        d_outs.append(dict(r=r.data_ptr(), J=J.data_ptr(), i0=idx[0].data_ptr(), i0_b=idx[1].data_ptr() if cam else None,
                           i0_c=idx[2].data_ptr() if cfg.get("split") else None, i0_d=idx[3].data_ptr() if (cfg.get("split") and cam) else None))

    dev_flags = flags | (_lib.EVAL_DEVICE_ORDER if a.row_order == "device" else 0)

    ne = None
    if strong:
        # the Levenberg-Marquardt linearisation on the device over the rows this rank holds (kontiki_b200/gn.py::DeviceSchurSolver, csrc/gn_device.cuh):
        # landmark blocks, diagonal knot blocks, gradient, reduced right-hand side, and one implicit-Schur product S p.  The exchange between
        # ranks is the all-reduce (NCCL, fp64 sum) of parameter-sized vectors: [c | g_rho | blocks | gradient] once, the reduced rhs once, S p once.
        import ctypes as C
        from kontiki_b200 import gn
        split = bool(cfg.get("split"))
        n_a, n_b = (len(cfg["r3"]), len(cfg["so3"])) if split else (len(cfg["knots"]), 0)
        GN_RADIUS = 1e4

        def make_ne(prob, c_):
            torch.cuda.set_stream(stream)
            hub = {g_: c_["cam"]["huber_c"] for g_ in range(prob.num_groups) if prob.group_kind(g_) == _lib.STATIC_RS}
            e = gn.DeviceSchurSolver(prob, split, n_a, n_b, 0 if rho is None else len(rho), local_rank, hubers=hub)
            e.set_point(knots_flat, rho)
            e.evaluate()                  # first evaluation: builds the row lists (ktk_gn_prepare)
            return e
        ne = make_ne(p, cfg)
        d_outs, keep = ne.outs, ne._keep
        dev_flags = ne.flags

        def gn_linearize(e):
            e.linearize(GN_RADIUS)
            e.p.gn_call("pcg_begin", C.c_double(GN_RADIUS), C.c_double(1e-6), C.c_int32(100))

        def gn_product(e):
            e.p.gn_call("product")
            e._reduce("qq")
            e.p.gn_call("pcg_update")

    def step_device():
        if ne is not None:
            ne.evaluate(cost=False)
            gn_linearize(ne)
            gn_product(ne)
            return
        p.evaluate_device(d_knots.data_ptr(), d_rho.data_ptr() if d_rho is not None else 0, 0 if rho is None else len(rho), dev_flags, d_outs)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step_device()
    p.synchronize()
    barrier()
    strong_graph = None
    if ne is not None and os.environ.get("KTK_STRONG_GRAPH", "1") != "0":
        # The strong-scaling step is ~25 short kernels, three all-reduces and a dozen torch ops: with the rows split over 8 GPUs the device work
        # shrinks to ~0.15 ms and the step was bound by the HOST issuing it (Python + ctypes launches, ~0.4 ms).  Every stretch between two
        # collectives is captured once in a CUDA graph (kernels of the library + the torch ops; the library's own per-evaluation graph is off
        # inside: graphs do not nest) and a step is 4 graph replays + 3 eager ncclAllReduce calls.  The collectives stay OUTSIDE the graphs.
        try:
            p.set_graphs(False)
            lc0 = p.launch_count
            step_device()
            torch.cuda.synchronize()
            launches_per_eager_step = p.launch_count - lc0      # kernels of the library in one step (the graphs replay exactly these)
            barrier()
            segments = [lambda: (ne.evaluate(cost=False), ne.linearize_local()),
                        lambda: ne.linearize_rhs(GN_RADIUS),
                        lambda: (ne.p.gn_call("pcg_begin", C.c_double(GN_RADIUS), C.c_double(1e-6), C.c_int32(100)), ne.p.gn_call("product")),
                        lambda: ne.p.gn_call("pcg_update")]
            graphs = []
            for seg in segments:
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=stream):
                    seg()
                graphs.append(gph)
            torch.cuda.synchronize()
            reduces = [("lin",), ("qq",), ("qq",)]
            eager_step = step_device

            def step_device():      # noqa: F811  (the timed loops below call this name)
                for k_, gph in enumerate(graphs):
                    gph.replay()
                    if k_ < 3:
                        ne._reduce(*reduces[k_])
            strong_graph = graphs
            for _ in range(3):
                step_device()
            torch.cuda.synchronize()
            barrier()
        except Exception as e:      # capture refused: keep the eager step, say so
            strong_graph = None
            sys.stderr.write(f"strong-scaling step not captured in CUDA graphs: {e}\n")
            p.set_graphs(True)
    l0 = p.launch_count
    # (1) the timed region: EXACTLY K steps between two events on the launching stream (each evaluation = one CUDA-graph launch), barrier +
    #     synchronize on both sides.  Nothing else runs on the host meanwhile: the nvidia-smi sampler (a fork per sample, and a driver query that
    #     every rank issues at once) is started AFTER the second event -- round 1 ran it inside the region.
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record(stream)
    for k in range(a.steps):
        step_device()
        ev[k + 1].record(stream)
    barrier()
    launches = p.launch_count - l0 if strong_graph is None else launches_per_eager_step * a.steps
    ms_total = ev[0].elapsed_time(ev[-1])
    per_step = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps)])
    with ClockSampler(local_rank) as clocks:
        # (2) the same steps again, the clocks sampled under this load: >= 200 steps and >= 0.5 s, one event pair per step (median / spread)
        n_steady = max(200, a.steps)
        sv = [torch.cuda.Event(enable_timing=True) for _ in range(n_steady + 1)]
        t_end = time.time() + (0.3 if a.quick else 0.6)
        sv[0].record(stream)
        for k in range(n_steady):
            step_device()
            sv[k + 1].record(stream)
        torch.cuda.synchronize()
        steady = np.array([sv[k].elapsed_time(sv[k + 1]) for k in range(n_steady)])
        while ne is None and time.time() < t_end:      # weak mode only: a time-based loop must not contain collectives (ranks would disagree on the count)
            step_device()
            torch.cuda.synchronize()
        # (3) every kernel launch bracketed by CUDA events inside the library (plain stream launches, no graph): per-kernel device time for the
        #     roofline of the dominant kernel
        p.set_profiling(True)
        for _ in range(min(a.steps, 50)):
            (eager_step if strong_graph is not None else step_device)()
        p.synchronize()
        prof = {name: p.read_profile(g) for name, g in groups.items()}
        p.set_profiling(False)
    ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    rank_ms = [ms.clone() for _ in range(world)]
    if dist is not None:
        dist.all_gather(rank_ms, ms)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / a.steps
    n_total = syn.num_measurements(cfg_full) if strong else world * n_meas
    value = n_total / (ms_step * 1e-3)
    timing = {"ms_per_step_median": float(np.median(per_step)), "ms_per_step_per_rank": [float(t.item()) / a.steps for t in rank_ms],
              "steady": {"steps": int(len(steady)), "ms_median": float(np.median(steady)), "ms_p10": float(np.percentile(steady, 10)), "ms_p90": float(np.percentile(steady, 90)),
                         "window_s": float(steady.sum() * 1e-3)}}

    strong_info = None
    if strong:
        # time of the pieces (max over ranks) and of the exchange alone; then the sharded linearisation against the UNSHARDED problem (rank 0)
        def timed(fn, reps=20):
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            b0.record(stream)
            for _ in range(reps):
                fn()
            b1.record(stream)
            barrier()
            t = torch.tensor([b0.elapsed_time(b1) / reps], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        # the pieces are timed EAGERLY (host-issued, one call after the other): their sum exceeds the graph-replayed step by the host's issue time
        parts = {"evaluate": timed(lambda: ne.evaluate(cost=False)), "linearize": timed(lambda: gn_linearize(ne)), "schur_product": timed(lambda: gn_product(ne))}
        ar = {"linearize": ("lin",), "rhs": ("qq",), "product": ("qq",)}
        ar_bytes = {k_: int(sum(ne.buf(n_).numel() for n_ in v_) * 8) for k_, v_ in ar.items()}
        ar_ms = {k_: (timed(lambda v_=v_: ne._reduce(*v_), 50) if world > 1 else 0.0) for k_, v_ in ar.items()}
        strong_info = {"cuda_graphs_between_collectives": strong_graph is not None,
                       "collective": ("ncclAllReduce(sum, fp64) of parameter-sized buffers: landmark blocks + gradient + diagonal knot blocks once per linearisation, "
                                      "the reduced right-hand side once, S p once per CG iteration; each exchange is ONE collective on a contiguous buffer") if world > 1 else "none (one rank)",
                       "allreduce_bytes": ar_bytes, "allreduce_ms": ar_ms, "ms": parts, "rows_this_rank": int(n_meas), "rows_total": int(n_total),
                       "step": "evaluation + LM linearisation (c, g_rho, B_kk, gradient, reduced rhs, block-Jacobi preconditioner) + one implicit-Schur product and CG update"}
        ne.evaluate(cost=False)
        gn_linearize(ne)
        gn_product(ne)
        got = {n_: ne.buf(n_).clone() for n_ in ("z_a", "z_b", "b_a", "b_b", "q_a", "q_b", "grho")}
        cost = ne.evaluate()
        if rank == 0:
            saved_dist = gn._dist
            gn._dist = lambda: None
            try:
                p0 = _lib.Problem(local_rank)
                p0.set_stream(stream.cuda_stream)
                if split:
                    p0.set_split_spline(cfg_full["dt"], cfg_full["t0"], n_a, cfg_full["dt"], cfg_full["t0"], n_b)
                else:
                    p0.set_se3_spline(cfg_full["dt"], cfg_full["t0"], n_a)
                for k_, add in (("gyro", p0.add_gyroscope), ("accel", p0.add_accelerometer)):
                    if cfg_full[k_]:
                        add(imu, cfg_full[k_]["t"], cfg_full[k_]["y"], cfg_full[k_]["weight"])
                if cfg_full["cam"]:
                    c = cfg_full["cam"]
                    p0.add_static_rs(_lib.make_camera(c["rows"], c["cols"], c["readout"], c["K"]), c["obs_uv"], c["obs_t0"], c["ref_uv"], c["ref_t0"], c["lm_idx"], c["weight"], c["huber_c"])
                ne0 = make_ne(p0, cfg_full)
                cost0 = ne0.evaluate()
                gn_linearize(ne0)
                gn_product(ne0)
                errs = {"rel_err_cost": abs(cost - cost0) / abs(cost0)}
                for n_ in got:
                    ref = ne0.buf(n_)
                    if ref.numel():
                        errs["rel_err_" + n_] = float((got[n_] - ref).abs().max() / ref.abs().max())
                errs["ok"] = bool(max(errs.values()) < 1e-10)
                strong_info["vs_unsharded"] = errs
                del ne0, p0
            finally:
                gn._dist = saved_dist

    # ---- end-to-end leg: host buffers through ktk_evaluate ---------------------------------------------------------
    h_knots = torch.from_numpy(knots_flat.copy()).pin_memory()
    h_rho = torch.from_numpy(rho).pin_memory() if rho is not None else None
    h_outs, d2h = [], 0
    for g in range(0 if a.quick else p.num_groups):
        n, cam = p.group_size(g), p.group_kind(g) in (_lib.STATIC_RS, _lib.NEWTON_RS, _lib.LIFTING_RS)
        o = dict(r=torch.empty((n, 2 if p.group_kind(g) in (_lib.STATIC_RS, _lib.NEWTON_RS) else 3), dtype=torch.float64).pin_memory(),
                 J=torch.empty((n, p.group_row_size(g)), dtype=torch.float64).pin_memory(),
                 i0=torch.empty(n, dtype=torch.int32).pin_memory())
        if cam:
            o["i0_b"] = torch.empty(n, dtype=torch.int32).pin_memory()
        if cfg.get("split"):
            o["i0_c"] = torch.empty(n, dtype=torch.int32).pin_memory()
            if cam:
                o["i0_d"] = torch.empty(n, dtype=torch.int32).pin_memory()
        d2h += sum(t.numel() * t.element_size() for t in o.values())
        h_outs.append({k: v.numpy() for k, v in o.items()})
        keep.append(o)
    h2d = h_knots.numel() * 8 + (h_rho.numel() * 8 if h_rho is not None else 0)
    e2e_steps = 0 if a.quick else max(3, min(a.steps, 10))
    def step_host():
        p.evaluate_flat(h_knots.numpy(), None if h_rho is None else h_rho.numpy(), flags, h_outs)

    for _ in range(0 if a.quick else 2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(e2e_s.item())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = "cam" if "cam" in groups else ("accel" if "accel" in groups else "gyro")
    dom_ms, dom_n = prof[dom]
    dom_rows = p.group_size(groups[dom])
    per_row = ({"cam": 1020, "accel": 744, "gyro": 452} if cfg.get("split") else {"cam": 1012, "accel": 740, "gyro": 740})[dom]
    if dom == "cam" and a.camera_method == "newton":      # in 76 + r 16 + packed row (58 + 14 W doubles) + i0_ref, i0_obs
        per_row = 76 + 16 + 8 * p.group_row_size(groups["cam"]) + 8
    if dom == "cam" and a.camera_method == "lifting":     # in 76 + vt 8 + r 24 + packed row (90 + 21 W doubles) + i0_ref, i0_obs
        per_row = 76 + 8 + 24 + 8 * p.group_row_size(groups["cam"]) + 8
    dom_bytes = dom_rows * per_row
    achieved = dom_bytes / (dom_ms / max(dom_n, 1) * 1e-3) / 1e9 if dom_ms > 0 else None
    traffic, traffic_src = None, None      # NOT measured by this run: dram bytes per launch of the same kernel from the committed ncu --set full capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get({"cam": "k_static_rs" if a.camera_method == "static" else "-", "accel": "k_imu<1>", "gyro": "k_imu<0>"}[dom])
        traffic_src = tj.get("_source")
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a.workload, cfg_full if strong else cfg, "device" if strong else a.row_order, a.camera_method, a.camera_model, strong, world),
            "timing": timing,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "ktk_evaluate (C ABI, pinned host buffers; every residual, Jacobian row and index copied back)", "numa": numa},
            "gpu_launches": int(launches),
            "clocks": dict(clocks.summary(), sampled="under the same steps, immediately after the timed region (sampler outside the event pair)"),
            "roofline": {"bound": "hbm", "kernel": {"cam": dict(newton="k_landmark_ref + k_newton_rs_fast + k_newton_rs_rev (+ k_newton_rs)", lifting="k_lifting_rs").get(a.camera_method, "k_static_rs"), "accel": "k_imu<1>", "gyro": "k_imu<0>"}[dom],
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms / max(dom_n, 1), "launches_timed": dom_n,
                         "kernel_ms_per_step": {k: v[0] / max(v[1], 1) for k, v in prof.items()},
                         # secondary, compute side (north_star: "fp64-pipe utilisation against chip peak"): fp64 instructions per row from the
                         # ncu instruction mix (profiles/r1h_ncu_summary.csv / README.md: k_static_rs 1410 DFMA + 594 DMUL + 206 DADD per row, plus
                         # k_landmark_ref 1565 + 548 + 249 per landmark record, 0.1 records per row on H1: 3620 + 393 flop), peak = the
                         # dependent-DFMA microbenchmark tools/fp64_microbench.cu on this pool's B200 (profiles/r1_fp64_microbench.txt)
                         "fp64": ({"flop_per_row": 4013, "achieved_tflops": dom_rows * 4013 / (dom_ms / max(dom_n, 1) * 1e-3) / 1e12, "peak_tflops": 32.8,
                                   "frac": dom_rows * 4013 / (dom_ms / max(dom_n, 1) * 1e-3) / 1e12 / 32.8} if (dom == "cam" and not cfg.get("split") and dom_ms > 0 and a.camera_method == "static") else None)}}
    if line["roofline"].get("fp64"):      # scalars next to the nested object (a parser that keeps only scalars of `roofline` still sees the compute side)
        line["roofline"]["fp64_frac"] = line["roofline"]["fp64"]["frac"]
        line["roofline"]["fp64_achieved_tflops"] = line["roofline"]["fp64"]["achieved_tflops"]
        line["roofline"]["fp64_peak_tflops"] = line["roofline"]["fp64"]["peak_tflops"]
    if strong_info is not None:
        line["strong"] = strong_info
    if not a.quick:      # rank 0 checks its own shard (every rank's shard has the same construction)
        line["parity"] = parity_gate(cfg, p, groups, keep, rho, dev_flags, bool(dev_flags & _lib.EVAL_DEVICE_ORDER))
    if not a.no_cpu_baseline and not a.quick:      # rank 0's host cores (the other ranks have finished their work by now)
        base, _, _ = cpu_baseline(cfg_full if strong else cfg, a.cpu_sample)
        line["cpu_baseline"] = base
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
