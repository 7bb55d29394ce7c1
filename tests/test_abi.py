"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/kontiki_b200.h declares,
refuses to compute without a GPU, and its host-side structure bookkeeping (spline_base.h:361-404 rule) matches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fixtures_ref as fx
from kontiki_b200 import _lib, synthetic as syn
from oracle import kto

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "kontiki_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ktk_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 19
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/kontiki_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTS)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(_lib.KontikiError) as e:
        _lib.Problem(0)
    assert e.value.code == _lib.ECUDA
    # a host-only handle answers structure queries but refuses to evaluate
    p = _lib.Problem(-1)
    p.set_se3_spline(0.1, 0.0, 20)
    p.add_gyroscope(_lib.make_sensor(), [0.5], [[0, 0, 0]])
    with pytest.raises(_lib.KontikiError) as e:
        p.evaluate(np.zeros((20, 7)))
    assert e.value.code == _lib.ECUDA


def test_structure_unlocked_time_offset_matches_oracle():
    """gyroscope_measurement.h:83-91: with the time offset unlocked the span is t -+ max_time_offset => more knot blocks."""
    knots = fx.smooth_se3_knots(80, 0.05)
    t = np.random.default_rng(1).uniform(0.2, 3.6, 200)
    p = _lib.Problem(-1)
    p.set_se3_spline(0.05, 0.0, 80)
    g = p.add_gyroscope(_lib.make_sensor(time_offset_locked=False, max_time_offset=0.1), t, np.zeros((200, 3)))
    ids, nids = p.get_structure(g, cap=12)
    o = kto.imu_residuals(kto.Traj(kto.SE3, 0.05, 0.0, knots), kto.Sensor(d_locked=False, max_time_offset=0.1), 0, t, np.zeros((200, 3)), jac_mode=0, cap=12)
    assert (ids == o["ids_a"]).all()
    assert nids.min() >= 8


def test_structure_imu_matches_oracle():
    knots = fx.smooth_se3_knots(50, 0.1)
    t = np.random.default_rng(0).uniform(0, 4.69, 300)
    p = _lib.Problem(-1)
    p.set_se3_spline(0.1, 0.0, 50)
    g = p.add_gyroscope(_lib.make_sensor(), t, np.zeros((300, 3)))
    ids, nids = p.get_structure(g, cap=4)
    o = kto.imu_residuals(kto.Traj(kto.SE3, 0.1, 0.0, knots), kto.Sensor(), 0, t, np.zeros((300, 3)), jac_mode=0)
    assert (nids == 4).all()
    assert (ids == o["ids_a"]).all()          # bit-exact block lists


def test_structure_out_of_range_raises_like_reference():
    p = _lib.Problem(-1)
    p.set_se3_spline(0.1, 0.0, 20)            # valid time [0, 1.7)
    g = p.add_gyroscope(_lib.make_sensor(), [1.71], [[0, 0, 0]])
    with pytest.raises(ValueError):           # std::range_error -> ValueError (trajectory_estimator.h:106-116)
        p.get_structure(g, cap=4)


@pytest.mark.parametrize("dt", [0.02, 0.05, 0.1])
def test_structure_static_rs_matches_oracle(dt):
    n_knots = 120
    knots = syn.smooth_se3_knots(n_knots, dt)
    cam = syn.make_static_rs(knots, dt, 40, obs_per_landmark=6, seed=7)
    p = _lib.Problem(-1)
    p.set_se3_spline(dt, 0.0, n_knots)
    g = p.add_static_rs(_lib.make_camera(cam["rows"], cam["cols"], cam["readout"], cam["K"]), cam["obs_uv"], cam["obs_t0"], cam["ref_uv"],
                        cam["ref_t0"], cam["lm_idx"])
    ids, nids = p.get_structure(g, cap=24)
    ocam = kto.Camera(cam["rows"], cam["cols"], cam["readout"], K=cam["K"])
    o = kto.static_rs_residuals(kto.Traj(kto.SE3, dt, 0.0, knots), ocam, cam["obs_uv"], cam["obs_t0"], cam["ref_uv"], cam["ref_t0"], cam["lm_idx"],
                                cam["rho"], jac_mode=0, cap=24)
    assert (ids == o["ids_a"]).all()
    assert (nids == (o["ids_a"] >= 0).sum(1)).all()
    assert len(set(nids.tolist())) > 1        # merged and split segment cases both occur


def test_structure_of_the_added_measurement_kinds_matches_oracle():
    """Host-only handles answer the structure queries of every kind: OrientationMeasurement / PositionMeasurement carry the 4 knots of their
    segment (orientation_measurement.h:66, position_measurement.h:66), LiftingRs / NewtonRs the two camera spans; row and residual sizes."""
    dt, n_knots = 0.05, 120
    knots = syn.smooth_se3_knots(n_knots, dt)
    traj = kto.Traj(kto.SE3, dt, 0.0, knots)
    rng = np.random.default_rng(2)
    t = rng.uniform(0.1, dt * (n_knots - 3) - 0.1, 50)
    p = _lib.Problem(-1)
    p.set_se3_spline(dt, 0.0, n_knots)
    go = p.add_orientation(t, np.tile([0.0, 0.0, 0.0, 1.0], (50, 1)))
    gp = p.add_position(t, np.zeros((50, 3)))
    o = kto.imu_residuals(traj, kto.Sensor(), 0, t, np.zeros((50, 3)), jac_mode=0)
    for g, kind, row in ((go, _lib.ORIENTATION, 28), (gp, _lib.POSITION, 84)):
        ids, nids = p.get_structure(g, cap=4)
        assert (ids == o["ids_a"]).all() and (nids == 4).all()
        assert p.group_kind(g) == kind and p.group_row_size(g) == row
    cam = syn.make_static_rs(knots, dt, 30, obs_per_landmark=5, seed=9)
    ccam = _lib.make_camera(cam["rows"], cam["cols"], cam["readout"], cam["K"])
    args = (cam["obs_uv"], cam["obs_t0"], cam["ref_uv"], cam["ref_t0"], cam["lm_idx"])
    gl, gn = p.add_lifting_rs(ccam, *args), p.add_newton_rs(ccam, *args)
    oc = kto.static_rs_residuals(traj, kto.Camera(cam["rows"], cam["cols"], cam["readout"], K=cam["K"]), *args, cam["rho"], jac_mode=0, cap=24)
    for g in (gl, gn):
        ids, _ = p.get_structure(g, cap=24)
        assert (ids == oc["ids_a"]).all()
    W = (p.group_row_size(gn) - 58) // 14
    assert p.group_kind(gl) == _lib.LIFTING_RS and p.group_row_size(gl) == 90 + 21 * W and W >= 4
    with pytest.raises(_lib.KontikiError) as e:
        p.evaluate(knots, cam["rho"])
    assert e.value.code == _lib.ECUDA
    with pytest.raises(ValueError):
        p.set_group_vt(gn, np.zeros(len(cam["lm_idx"])))          # only LiftingRs groups have row-time parameters
