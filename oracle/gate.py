"""The sampled parity gate: rows produced by the CUDA path (packed layouts of include/kontiki_b200.h) against the CPU oracle.

TEST INFRASTRUCTURE ONLY -- the checker behind bench.py's `parity` key (SURVEY.md section 8d "Parity gate run with every benchmark") and
tests/test_gpu_fullsize.py.  It never produces anything the product ships or the benchmark times.
Bar (BASELINE.json north_star): knot / landmark indices bit-exact; residuals and Jacobians within 1e-9 relative to the largest entry of the
row; camera residuals are pixel differences of size ~1e3 and are compared absolutely, 1e-9 px."""
import numpy as np

from . import kto

TOL = 1e-9
CAM_R_TOL = 1e-9


def _relmax(a, b):
    a, b = np.asarray(a).reshape(len(a), -1), np.asarray(b).reshape(len(b), -1)
    return float((np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-300)).max()) if len(a) else 0.0


def _onto_ids(blocks, ids, flag):
    """[(packed (n, nk, nres, w), first knot (n,)), ...] -> (n, cap, nres, w) laid out like the oracle's structural block list `ids`."""
    n, _, nres, w = blocks[0][0].shape
    d = np.zeros((n, ids.shape[1], nres, w))
    rows = np.arange(n)
    for blk, st in blocks:
        for k in range(blk.shape[1]):
            hit = ids == (st + k)[:, None]
            if not hit.any(1).all():
                flag[0] = False           # an active knot that is not in the reference's block list
            d[rows, hit.argmax(1)] += blk[:, k] * hit.any(1)[:, None, None]
    return d


def make_traj(cfg):
    if cfg.get("split"):
        return kto.Traj(kto.SPLIT, cfg["dt"], cfg["t0"], cfg["r3"], cfg["dt"], cfg["t0"], cfg["so3"])
    return kto.Traj(kto.SE3, cfg["dt"], cfg["t0"], cfg["knots"])


def check_span_rows(cfg, sel, r, J, idx, rho=None, robust=False, atan=None, method="newton", vt=None, nthreads=0):
    """NewtonRs / LiftingRs rows on an SE3 trajectory: packed [ref 4 x (nres x 7) | obs W x (nres x 7) | (vt nres) | rho nres], idx = [i0_ref, first knot
    of the observation span].  The oracle differentiates through the Newton iteration (ceres::Jet semantics); the rows are laid onto its structural
    block list, whole rows are compared (the Huber corrector acts on the whole row)."""
    traj = make_traj(cfg)
    c = cfg["cam"]
    ns = len(sel)
    nres = 2 if method == "newton" else 3
    tail = 2 if method == "newton" else 6
    J = np.asarray(J).reshape(ns, -1)
    W = (J.shape[1] - tail) // (7 * nres) - 4
    ok = [(J.shape[1] - tail) % (7 * nres) == 0 and W >= 4]
    ocam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"], q_ct=c.get("q_ct", (0, 0, 0, 1)), p_ct=c.get("p_ct", (0, 0, 0)),
                      method="newton" if method == "newton" else "static", **(atan or {}))
    args = (c["obs_uv"][sel], c["obs_t0"][sel], c["ref_uv"][sel], c["ref_t0"][sel], c["lm_idx"][sel], c["rho"] if rho is None else rho)
    if method == "newton":
        o = kto.static_rs_residuals(traj, ocam, *args, c["weight"][sel], jac_mode=2, cap=8 + W, nthreads=nthreads)
    else:
        o = kto.lifting_rs_residuals(traj, ocam, *args, vt=None if vt is None else np.asarray(vt)[sel], weight=c["weight"][sel], jac_mode=2, cap=8 + W, nthreads=nthreads)
    ok[0] &= bool(np.array_equal(idx[0], o["i0_ref_a"]))
    if method == "newton":      # the oracle reports the first knot of the observation span for Newton rows, the active window of the lifted time for lifting rows
        ok[0] &= bool(np.array_equal(idx[1], o["i0_obs_a"]))
    else:
        ok[0] &= bool(((o["i0_obs_a"] >= idx[1]) & (o["i0_obs_a"] + 4 <= np.asarray(idx[1]) + W)).all())
    # the span's trailing blocks may reach past the last knot of the spline for rows at its very end: structurally absent there, and zero in the row
    n_knots = len(cfg["knots"])
    obs = J[:, 28 * nres:28 * nres + 7 * nres * W].reshape(ns, W, nres, 7).copy()
    past = (np.asarray(idx[1])[:, None] + np.arange(W)[None, :]) >= n_knots
    ok[0] &= not obs[past].any()
    obs[past] = 0.0
    ids = o["ids_a"].copy()
    mine = np.zeros((ns, ids.shape[1], nres, 7))
    rows = np.arange(ns)
    for blk, st in ((J[:, :28 * nres].reshape(ns, 4, nres, 7), np.asarray(idx[0])), (obs, np.asarray(idx[1]))):
        for k in range(blk.shape[1]):
            hit = ids == (st + k)[:, None]
            found = hit.any(1)
            # a block outside the reference's list for this row (the packed row carries the widest span of the group) must be structurally zero
            if blk[~found, k].any():
                ok[0] = False
            mine[rows, hit.argmax(1)] += blk[:, k] * found[:, None, None]
    tails = J[:, -tail:].reshape(ns, tail // nres, nres).transpose(0, 2, 1)              # (ns, nres, 1 or 2): [vt |] rho columns
    ref_tails = o["Jrho"].reshape(ns, nres, 1) if method == "newton" else np.stack([o["Jvt"], o["Jrho"]], axis=2)
    rows_mine = np.concatenate([mine.transpose(0, 2, 1, 3).reshape(ns, nres, -1), tails], axis=2)
    rows_ref = np.concatenate([o["Ja"].transpose(0, 2, 1, 3).reshape(ns, nres, -1), ref_tails], axis=2)
    r_ref = o["r"].copy()
    if robust:
        for i in range(ns):
            a = float(c["huber_c"][sel[i]])
            if float(r_ref[i] @ r_ref[i]) > a * a:
                _, r_ref[i], rows_ref[i] = kto.huber_correct(a, r_ref[i], rows_ref[i])
    return dict(idx_exact=bool(ok[0]), rel_r=0.0, rel_J=_relmax(rows_mine, rows_ref), abs_r_cam_px=float(np.abs(np.asarray(r) - r_ref).max()) if ns else 0.0)


def check_rows(cfg, name, sel, r, J, idx, rho=None, robust=False, atan=None, nthreads=0, method="static", vt=None):
    """cfg: kontiki_b200.synthetic workload dict; name: "gyro" | "accel" | "cam"; sel: caller-order indices of the rows in (r, J, idx);
    idx: [i0, i0_b, i0_c, i0_d] as the C ABI writes them.  Returns dict(idx_exact, rel_r, rel_J, abs_r_cam_px)."""
    if name == "cam" and method != "static":
        return check_span_rows(cfg, sel, r, J, idx, rho, robust, atan, method, vt, nthreads)
    split = bool(cfg.get("split"))
    traj = make_traj(cfg)
    ns = len(sel)
    ok = [True]
    res = dict(idx_exact=True, rel_r=0.0, rel_J=0.0, abs_r_cam_px=0.0)
    J = np.asarray(J).reshape(ns, -1)
    if name in ("gyro", "accel"):
        which = 0 if name == "gyro" else 1
        m = cfg[name]
        o = kto.imu_residuals(traj, kto.Sensor(), which, m["t"][sel], m["y"][sel], m["weight"][sel], jac_mode=2, nthreads=nthreads)
        if not split:
            ok[0] &= bool(np.array_equal(idx[0], o["i0_a"]))
            res["rel_J"] = _relmax(J.reshape(ns, 4, 3, 7), o["Ja"][:, :4])
        elif which == 0:
            ok[0] &= bool(np.array_equal(idx[2], o["i0_b"]))
            res["rel_J"] = _relmax(J.reshape(ns, 4, 3, 4), o["Jb"][:, :4])
        else:
            ok[0] &= bool(np.array_equal(idx[0], o["i0_a"]) and np.array_equal(idx[2], o["i0_b"]))
            res["rel_J"] = max(_relmax(J[:, :36].reshape(ns, 4, 3, 3), o["Ja"][:, :4]), _relmax(J[:, 36:].reshape(ns, 4, 3, 4), o["Jb"][:, :4]))
        res["rel_r"] = _relmax(r, o["r"])
    else:
        c = cfg["cam"]
        ocam = kto.Camera(c["rows"], c["cols"], c["readout"], K=c["K"], q_ct=c.get("q_ct", (0, 0, 0, 1)), p_ct=c.get("p_ct", (0, 0, 0)), **(atan or {}))
        o = kto.static_rs_residuals(traj, ocam, c["obs_uv"][sel], c["obs_t0"][sel], c["ref_uv"][sel], c["ref_t0"][sel], c["lm_idx"][sel], c["rho"] if rho is None else rho,
                                    c["weight"][sel], jac_mode=2, cap=24, nthreads=nthreads)
        if not split:
            ok[0] &= bool(np.array_equal(idx[0], o["i0_ref_a"]) and np.array_equal(idx[1], o["i0_obs_a"]))
            mine = [_onto_ids([(J[:, :56].reshape(ns, 4, 2, 7), idx[0]), (J[:, 56:112].reshape(ns, 4, 2, 7), idx[1])], o["ids_a"], ok)]
            ref = [o["Ja"]]
        else:
            ok[0] &= bool(np.array_equal(idx[0], o["i0_ref_a"]) and np.array_equal(idx[1], o["i0_obs_a"]) and np.array_equal(idx[2], o["i0_ref_b"])
                          and np.array_equal(idx[3], o["i0_obs_b"]))
            mine = [_onto_ids([(J[:, 0:24].reshape(ns, 4, 2, 3), idx[0]), (J[:, 56:80].reshape(ns, 4, 2, 3), idx[1])], o["ids_a"], ok),
                    _onto_ids([(J[:, 24:56].reshape(ns, 4, 2, 4), idx[2]), (J[:, 80:112].reshape(ns, 4, 2, 4), idx[3])], o["ids_b"], ok)]
            ref = [o["Ja"], o["Jb"]]
        # one row = [all knot columns | rho]; the Huber corrector acts on the whole row (Ceres applies it after Evaluate)
        rows_mine = np.concatenate([m_.transpose(0, 2, 1, 3).reshape(ns, 2, -1) for m_ in mine] + [J[:, 112:114].reshape(ns, 2, 1)], axis=2)
        rows_ref = np.concatenate([m_.transpose(0, 2, 1, 3).reshape(ns, 2, -1) for m_ in ref] + [o["Jrho"].reshape(ns, 2, 1)], axis=2)
        r_ref = o["r"].copy()
        if robust:
            for i in range(ns):
                a = float(c["huber_c"][sel[i]])
                if float(r_ref[i] @ r_ref[i]) > a * a:
                    _, r_ref[i], rows_ref[i] = kto.huber_correct(a, r_ref[i], rows_ref[i])
        res["abs_r_cam_px"] = float(np.abs(np.asarray(r) - r_ref).max()) if ns else 0.0
        res["rel_J"] = _relmax(rows_mine, rows_ref)
    res["idx_exact"] = bool(ok[0])
    return res
