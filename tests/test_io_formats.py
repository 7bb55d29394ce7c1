"""Host-side formats: structure / trajectory files in the reference's layout (python/kontiki/io.py) and the
sfm object graph <-> flat index arrays round trip.  No GPU needed."""
import numpy as np
import pytest

import fixtures_ref as fx
from kontiki_b200 import io, sfm
from kontiki_b200.trajectories import SplitTrajectory, UniformR3SplineTrajectory, UniformSE3SplineTrajectory, UniformSO3SplineTrajectory


def _structure(n_lm=12, n_views=7, seed=0):
    rng = np.random.default_rng(seed)
    views = [sfm.View(10 + i, i / 30) for i in range(n_views)]
    landmarks = []
    for _ in range(n_lm):
        lm = sfm.Landmark()
        seen = sorted(rng.choice(n_views, size=rng.integers(2, n_views), replace=False))
        obs = [views[v].create_observation(lm, rng.uniform(0, 1000, 2)) for v in seen]
        lm.reference = obs[rng.integers(0, len(obs))]          # not necessarily the first observation
        lm.inverse_depth = rng.uniform(0.01, 2)
        lm.locked = bool(rng.integers(0, 2))
        landmarks.append(lm)
    return views, landmarks


def _same_structure(a, b):
    assert len(a) == len(b)
    for la, lb in zip(a, b):
        assert la.inverse_depth == lb.inverse_depth and len(la.observations) == len(lb.observations)
        assert la.reference.view.frame_nr == lb.reference.view.frame_nr and np.array_equal(la.reference.uv, lb.reference.uv)
        for oa, ob in zip(la.observations, lb.observations):
            assert np.array_equal(oa.uv, ob.uv) and oa.view.t0 == ob.view.t0 and oa.view.frame_nr == ob.view.frame_nr
            assert oa.is_reference == ob.is_reference


@pytest.mark.parametrize("colors", [False, True])
def test_structure_file_round_trip(tmp_path, colors):
    views, landmarks = _structure()
    cmap = {lm: np.random.default_rng(lm.id).integers(0, 255, 3) for lm in landmarks} if colors else None
    path = tmp_path / "structure.npz"
    io.save_structure(path, landmarks, landmark_colors=cmap)
    views2, landmarks2, cmap2 = io.load_structure(path)
    _same_structure(landmarks, landmarks2)
    used = sorted({o.view.frame_nr for lm in landmarks for o in lm.observations})
    assert [v.frame_nr for v in views2] == used                # io.py:143-144: views sorted by frame number
    assert (cmap2 is None) == (not colors)
    if colors:
        assert all(np.array_equal(cmap[a], cmap2[b]) for a, b in zip(landmarks, landmarks2))
    with np.load(path) as z:                                   # the reference's dataset paths (io.py:154-167)
        assert {"structure/views/frame_nr", "structure/views/t0", "structure/landmarks/inverse_depth", "structure/landmarks/ref_idx",
                "structure/landmarks/color", "structure/observations/uv", "structure/observations/lm_idx", "structure/observations/v_idx"} <= set(z.files)


def _trajectories():
    se3 = UniformSE3SplineTrajectory(fx.SE3_DT, fx.SE3_T0)
    for cp in fx.SE3_KNOTS:
        T = np.eye(4)
        T[:3, :3] = fx.rot_xyzw(cp[:4])
        T[:3, 3] = cp[4:7]
        se3.append_knot(T)
    r3 = UniformR3SplineTrajectory(fx.R3_DT, fx.R3_T0)
    for cp in fx.R3_KNOTS:
        r3.append_knot(cp)
    so3 = UniformSO3SplineTrajectory(fx.SO3_DT, fx.SO3_T0)
    for q in fx.SO3_KNOTS:
        so3.append_knot(fx.xyzw_to_wxyz(q))
    return [se3, r3, so3, SplitTrajectory(r3.clone(), so3.clone())]


@pytest.mark.parametrize("k", range(4))
def test_trajectory_file_round_trip(tmp_path, k):
    traj = _trajectories()[k]
    path = tmp_path / "traj.npz"
    io.save_trajectory(path, traj)
    back = io.load_trajectory(path)
    assert type(back) is type(traj)
    pairs = [(traj, back)] if not isinstance(traj, SplitTrajectory) else [(traj.R3_spline, back.R3_spline), (traj.SO3_spline, back.SO3_spline)]
    for a, b in pairs:
        assert a.dt == b.dt and a.t0 == b.t0 and len(a) == len(b)
        # SE3 knots go through 4x4 matrices: "precision loss in the order of 1e-16" (io.py:73-74)
        assert np.abs(a.control_points - b.control_points).max() < (1e-14 if isinstance(a, UniformSE3SplineTrajectory) else 1e-300)


def test_two_objects_in_one_container(tmp_path):
    """location may be an open group: structure and trajectory side by side, as the reference's scripts store them in one HDF5 file."""
    views, landmarks = _structure(5, 4, seed=3)
    store = {}
    root = io._NpzGroup(store, "")
    io.save_structure(root, landmarks)
    io.save_trajectory(root, _trajectories()[3])
    views2, lm2, _ = io.load_structure(root)      # the views own the observations (as in the reference): keep them
    _same_structure(landmarks, lm2)
    assert isinstance(io.load_trajectory(root), SplitTrajectory)


def test_flatten_structure_round_trip():
    views, landmarks = _structure(20, 9, seed=5)
    flat = io.flatten_structure(landmarks)
    n = sum(len(lm.observations) - 1 for lm in landmarks)
    assert flat["obs_uv"].shape == (n, 2) and flat["lm_idx"].dtype == np.int32 and flat["rho"].shape == (len(landmarks),)
    # every row points at its landmark's reference observation (static_rscamera_measurement.h:89-94)
    for i in range(n):
        lm = landmarks[flat["lm_idx"][i]]
        assert np.array_equal(flat["ref_uv"][i], lm.reference.uv) and flat["ref_t0"][i] == lm.reference.view.t0
    views2, landmarks2 = io.structure_from_arrays(**flat)
    flat2 = io.flatten_structure(landmarks2)
    for k in flat:
        assert np.array_equal(flat[k], flat2[k]), k
    assert [v.frame_nr for v in views2] == sorted({o.view.frame_nr for lm in landmarks for o in lm.observations})
    # without frame numbers the views are identified by their t0
    views3, landmarks3 = io.structure_from_arrays(flat["obs_uv"], flat["obs_t0"], flat["ref_uv"], flat["ref_t0"], flat["lm_idx"], flat["rho"])
    flat3 = io.flatten_structure(landmarks3)
    for k in ("obs_uv", "obs_t0", "ref_uv", "ref_t0", "lm_idx", "rho"):
        assert np.array_equal(flat[k], flat3[k]), k


def test_load_atan_camera(tmp_path):
    path = tmp_path / "cam.npz"
    K = np.array([[853.1, 0, 988.0], [0, 873.5, 525.7], [0, 0, 1]])
    with open(path, "wb") as f:
        np.savez(f, size=np.array([1920, 1080]), readout=0.026, K=K, wc=np.array([0.003, 0.0004]), lgamma=0.889)
    cam = io.load_atan_camera(path)
    assert (cam.rows, cam.cols, cam.readout, cam.gamma) == (1080, 1920, 0.026, 0.889) and np.array_equal(cam.camera_matrix, K)
