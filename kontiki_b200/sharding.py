"""Sharding of a measurement set across GPUs.

Measurements are independent units: every rank keeps a replica of the knots and of the landmark inverse depths and a
disjoint subset of the rows; residual + Jacobian evaluation needs no exchange at all.  The partition keeps
  * IMU rows in contiguous time ranges (rows sorted by time => each rank touches one contiguous knot window), and
  * all observations of a landmark on ONE rank (so the hoisted landmark-reference record and, later, the per-landmark
    Schur elimination of rho are rank-local).
Pure host logic (numpy); the multi-process tests run it under torch.distributed/gloo on CPU.
"""
import numpy as np


def shard_imu(t, rank, world):
    """Indices (into the caller's arrays, ascending time) of the rows rank `rank` of `world` owns: equal-count time ranges."""
    t = np.asarray(t)
    order = np.argsort(t, kind="stable")
    bounds = np.linspace(0, len(t), world + 1).astype(np.int64)
    return np.sort(order[bounds[rank]:bounds[rank + 1]])


def shard_static_rs(lm_idx, ref_t0, rank, world):
    """Rows of rank `rank`: landmarks are ordered by the time of their reference view and cut into `world` groups with
    (nearly) equal numbers of OBSERVATIONS; every observation of a landmark goes with its landmark."""
    lm_idx = np.asarray(lm_idx)
    ref_t0 = np.asarray(ref_t0)
    if len(lm_idx) == 0:
        return np.zeros(0, np.int64)
    n_lm = int(lm_idx.max()) + 1
    count = np.bincount(lm_idx, minlength=n_lm)
    t_lm = np.full(n_lm, np.inf)
    np.minimum.at(t_lm, lm_idx, ref_t0)
    order = np.argsort(t_lm, kind="stable")                 # landmarks by reference time (unused ids sort last, count 0)
    csum = np.cumsum(count[order])
    total = csum[-1]
    owner_sorted = np.minimum((csum - 1) * world // max(total, 1), world - 1)       # group of each landmark, in sorted order
    owner = np.empty(n_lm, np.int64)
    owner[order] = owner_sorted
    return np.nonzero(owner[lm_idx] == rank)[0]


def shard_config(cfg, rank, world):
    """Shard of a kontiki_b200.synthetic config dict (same keys; rho / knots replicated)."""
    out = dict(cfg)
    for k in ("gyro", "accel"):
        if cfg.get(k):
            sel = shard_imu(cfg[k]["t"], rank, world)
            out[k] = {a: v[sel] for a, v in cfg[k].items()}
            out[k]["rows"] = sel
    if cfg.get("cam"):
        c = cfg["cam"]
        sel = shard_static_rs(c["lm_idx"], c["ref_t0"], rank, world)
        out["cam"] = dict(c, **{a: c[a][sel] for a in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "weight", "huber_c")})
        out["cam"]["rows_sel"] = sel
    return out
