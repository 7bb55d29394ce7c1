"""The Gauss-Newton / LM step on the device (csrc/gn_device.cuh, ktk_gn_* of the C ABI, kontiki_b200/gn.py::DeviceSchurSolver) against
scipy on the host-assembled local Jacobian: landmark blocks, diagonal knot blocks, gradient, reduced right-hand side, the implicit Schur
product, the CG solution, the step of the eliminated inverse depths, the model decrease and the retraction -- and bit-for-bit
reproducibility (no atomics anywhere).  What the reference delegates to ceres::Solve / SPARSE_SCHUR (trajectory_estimator.h:38-64)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import kontiki_b200 as kontiki
from kontiki_b200 import gn
from test_python_surface import _vi_problem

pytestmark = pytest.mark.gpu
RADIUS = 1e3


def _setup(split, lock_some=False):
    start, ms, lms = _vi_problem(split)
    if lock_some:
        for L in lms[::5]:
            L.locked = True
    est = kontiki.TrajectoryEstimator(start)
    for m in ms:
        est.add_measurement(m)
    outs = est.evaluate(jacobians=True)
    r, J, layout = est._sparse_system(outs)                 # local coordinates, free columns only
    if split:
        a, b = start.R3_spline, start.SO3_spline
        kf, n_a, n_b = np.concatenate([a.control_points.reshape(-1), b.control_points.reshape(-1)]), len(a), len(b)
    else:
        kf, n_a, n_b = start.control_points.reshape(-1), len(start), 0
    locked = np.array([L.locked for L in lms], np.uint8)
    hubers = {grp["g"]: grp["huber"] for grp in est._groups if grp["kind"] == "cam"}
    sol = gn.DeviceSchurSolver(est._problem, split, n_a, n_b, len(lms), 0, locked, False, False, hubers)
    sol.set_point(kf, np.array([L.inverse_depth for L in lms]))
    return est, sol, r, J, lms, n_a, n_b, locked.astype(bool)


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("lock_some", [False, True])
def test_device_step_matches_scipy(split, lock_some):
    est, sol, r, J, lms, n_a, n_b, locked = _setup(split, lock_some)
    nk = (3 * n_a + 3 * n_b) if split else 6 * n_a
    cost = sol.evaluate()
    assert np.isclose(cost, est._cost(est.evaluate(), est._groups), rtol=1e-12)
    Jk, Jr = J[:, :nk].tocsc(), J[:, nk:].tocsc()            # knot columns, free rho columns
    free_idx = np.nonzero(~locked)[0]
    gmax = float(sol.linearize(RADIUS).item())
    g = J.T @ r
    assert np.isclose(gmax, np.abs(g).max(), rtol=1e-9)
    # landmark blocks and gradient
    c_ref = np.asarray(Jr.multiply(Jr).sum(0)).reshape(-1)
    assert _rel(sol.buf("c").cpu().numpy()[free_idx], c_ref) < 1e-12
    assert _rel(sol.buf("grho").cpu().numpy()[free_idx], Jr.T @ r) < 1e-10
    gk = np.concatenate([sol.buf("z_a").cpu().numpy(), sol.buf("z_b").cpu().numpy()])
    assert _rel(gk, Jk.T @ r) < 1e-10
    # exact diagonal knot blocks of J^T J (a knot in both windows of a camera row is ONE column block)
    B = (Jk.T @ Jk).tocsr()
    lw = 3 if split else 6
    blocks = np.concatenate([sol.buf("blocks_a").cpu().numpy().reshape(n_a, lw, lw)] + ([sol.buf("blocks_b").cpu().numpy().reshape(n_b, 3, 3)] if split else []))
    ref = np.stack([B[lw * k:lw * k + lw, lw * k:lw * k + lw].toarray() for k in range(nk // lw)])
    assert _rel(blocks, ref) < 1e-10
    # reduced right-hand side
    cd = c_ref + np.clip(c_ref, 1e-6, 1e32) / RADIUS
    E = (Jk.T @ Jr).tocsc()
    b_ref = -(Jk.T @ r - E @ ((Jr.T @ r) / cd))
    import ctypes as C
    sol.p.gn_call("pcg_begin", C.c_double(RADIUS), C.c_double(1e-12), C.c_int32(500))
    b_dev = np.concatenate([sol.buf("b_a").cpu().numpy(), sol.buf("b_b").cpu().numpy()])
    assert _rel(b_dev, b_ref) < 1e-9
    # implicit Schur product on a random vector (without the damping term, which pcg_update adds)
    v = np.random.default_rng(0).normal(size=nk)
    pa, pb = sol.buf("p_a"), sol.buf("p_b")
    import torch
    pa.copy_(torch.from_numpy(v[:pa.numel()]).to(sol.dev))
    if pb.numel():
        pb.copy_(torch.from_numpy(v[pa.numel():]).to(sol.dev))
    sol.p.gn_call("product")
    q = np.concatenate([sol.buf("q_a").cpu().numpy(), sol.buf("q_b").cpu().numpy()])
    S_v = B @ v - E @ ((E.T @ v) / cd)
    assert _rel(q, S_v) < 1e-9
    # the CG solution of the damped reduced system, delta_rho, model decrease, retraction
    sol.linearize(RADIUS)                                   # the product above overwrote the reduced gradient
    it, rel = sol.solve(RADIUS, tol=1e-12, max_iter=500)
    Dk = np.clip(B.diagonal(), 1e-6, 1e32) / RADIUS
    S = (B - E @ sp.diags(1.0 / cd) @ E.T + sp.diags(Dk)).tocsc()
    x_ref = spla.spsolve(S, b_ref)
    x = np.concatenate([sol.buf("x_a").cpu().numpy(), sol.buf("x_b").cpu().numpy()])
    assert rel < 1e-11 and _rel(x, x_ref) < 1e-7
    model, step = sol.finish()
    drho_ref = -((Jr.T @ r) + E.T @ x) / cd
    drho = sol.buf("drho").cpu().numpy()
    assert _rel(drho[free_idx], drho_ref) < 1e-7 and not drho[locked].any()
    delta = np.concatenate([x, drho[free_idx]])
    model_ref = -float(delta @ (g + 0.5 * (J.T @ (J @ delta))))
    assert np.isclose(model, model_ref, rtol=1e-8) and np.isclose(step, np.linalg.norm(delta), rtol=1e-10)
    # Plus() on the device == the host retraction
    from kontiki_b200.estimator import _quat_plus, _se3_plus
    kf_new = sol.knots_new.cpu().numpy()
    kf_old = sol.knots.cpu().numpy()
    if split:
        assert _rel(kf_new[:3 * n_a], kf_old[:3 * n_a] + x[:3 * n_a]) < 1e-14
        assert _rel(kf_new[3 * n_a:].reshape(n_b, 4), _quat_plus(kf_old[3 * n_a:].reshape(n_b, 4), x[3 * n_a:].reshape(n_b, 3))) < 1e-13
    else:
        assert _rel(kf_new.reshape(n_a, 7), _se3_plus(kf_old.reshape(n_a, 7), x.reshape(n_a, 6))) < 1e-13
    rho_old = sol.rho.cpu().numpy()
    assert np.array_equal(sol.rho_new.cpu().numpy(), np.maximum(0.0, rho_old + drho))
    est._problem.set_stream(0)


@pytest.mark.parametrize("split", [False, True])
def test_device_step_is_bit_reproducible(split):
    """Two independent runs of evaluation + linearisation + 25 CG iterations + finish give identical bits (fixed-order gathers, no atomics)."""
    res = []
    for _ in range(2):
        est, sol, r, J, lms, n_a, n_b, locked = _setup(split)
        sol.evaluate()
        sol.linearize(RADIUS)
        sol.solve(RADIUS, tol=0.0, max_iter=24, check_every=8)
        model, step = sol.finish()
        res.append((np.concatenate([sol.buf("x_a").cpu().numpy(), sol.buf("x_b").cpu().numpy()]), sol.buf("drho").cpu().numpy().copy(), sol.knots_new.cpu().numpy(), model, step))
        est._problem.set_stream(0)
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    assert np.abs(res[0][0]).max() > 0
