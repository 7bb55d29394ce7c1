"""Host-side formats: the reference's structure / trajectory files and the flat index arrays the device uses.

`save_structure / load_structure / save_trajectory / load_trajectory / load_atan_camera` keep the names, arguments and the dataset
layout of python/kontiki/io.py:13-103 (groups `views`, `landmarks`, `observations`; `type`, `dt`, `t0`, `knots`).  The reference
stores them in HDF5; h5py is used when it is importable, otherwise the SAME dataset paths are written into a NumPy `.npz` archive
(`structure/views/t0`, ...), so files round-trip in either environment and an HDF5 file written by the reference loads unchanged
where h5py exists.

`flatten_structure / structure_from_arrays` are the two directions of SURVEY.md 8a row a18: the sfm object graph
(sfm/landmark.h:19-54, observation.h:13-35, view.h:17-35) <-> the arrays of ktk_add_static_rs / ktk_add_newton_rs
(obs_uv, obs_t0, ref_uv, ref_t0, lm_idx) plus rho.
"""
import numpy as np

from .sensors import AtanCamera
from .sfm import Landmark, View
from .trajectories import SplitTrajectory, UniformR3SplineTrajectory, UniformSE3SplineTrajectory, UniformSO3SplineTrajectory

try:                                      # the reference's container
    import h5py
except ImportError:                       # not installed here: same dataset paths in an .npz archive
    h5py = None

_SPLINES = {c.__name__: c for c in (UniformSE3SplineTrajectory, UniformR3SplineTrajectory, UniformSO3SplineTrajectory)}


# ---- a minimal group interface over the two containers ----------------------------------------------------------------------
class _NpzGroup:
    """`group[name] = array`, `group[name]`, `create_group(name)` on a flat dict of 'a/b/c' keys."""

    def __init__(self, store, prefix):
        self._store, self._prefix = store, prefix

    def _key(self, name):
        return f"{self._prefix}/{name}" if self._prefix else name

    def create_group(self, name):
        return _NpzGroup(self._store, self._key(name))

    def __setitem__(self, name, value):
        self._store[self._key(name)] = np.asarray(value)

    def __getitem__(self, name):
        key = self._key(name)
        if key in self._store:
            return self._store[key]
        if any(k.startswith(key + "/") for k in self._store):
            return _NpzGroup(self._store, key)
        raise KeyError(key)

    def __contains__(self, name):
        key = self._key(name)
        return key in self._store or any(k.startswith(key + "/") for k in self._store)


def _value(ds):
    """Dataset -> numpy value (h5py datasets: `[()]`; npz: already arrays)."""
    v = ds[()] if hasattr(ds, "shape") and not isinstance(ds, np.ndarray) else ds
    if isinstance(v, bytes):
        v = v.decode()
    if isinstance(v, np.ndarray) and v.shape == () and v.dtype.kind in "US":
        v = str(v)
    return v


def _is_group(obj):
    return isinstance(obj, _NpzGroup) or (h5py is not None and isinstance(obj, (h5py.File, h5py.Group)))


def _use_h5(path):
    return h5py is not None and not str(path).endswith(".npz")


class _Writer:
    def __init__(self, location, group_name):
        self.location, self.group_name, self.file, self.store = location, group_name, None, None

    def __enter__(self):
        if _is_group(self.location):
            return self.location.create_group(self.group_name)
        if _use_h5(self.location):
            self.file = h5py.File(self.location, "w")
            return self.file.create_group(self.group_name)
        self.store = {}
        return _NpzGroup(self.store, self.group_name)

    def __exit__(self, *exc):
        if self.file is not None:
            self.file.close()
        elif self.store is not None and exc[0] is None:
            path = str(self.location)
            with open(path, "wb") as f:       # np.savez appends '.npz' to bare names: write through a handle to keep the caller's path
                np.savez(f, **self.store)


class _Reader:
    def __init__(self, location, group_name):
        self.location, self.group_name, self.file = location, group_name, None

    def __enter__(self):
        if _is_group(self.location):
            return self.location[self.group_name]
        path = str(self.location)
        with open(path, "rb") as f:
            magic = f.read(4)
        if magic[:2] == b"PK":                # zip container = npz
            with np.load(path, allow_pickle=False) as z:
                store = {k: z[k] for k in z.files}
            return _NpzGroup(store, "")[self.group_name]
        if h5py is None:
            raise IOError(f"{path} is an HDF5 file and h5py is not installed")
        self.file = h5py.File(path, "r")
        return self.file[self.group_name]

    def __exit__(self, *exc):
        if self.file is not None:
            self.file.close()


class _RootReader(_Reader):
    def __init__(self, location):
        super().__init__(location, None)

    def __enter__(self):
        path = str(self.location)
        with open(path, "rb") as fh:
            magic = fh.read(4)
        if magic[:2] == b"PK":
            with np.load(path, allow_pickle=False) as z:
                return _NpzGroup({k: z[k] for k in z.files}, "")
        if h5py is None:
            raise IOError(f"{path} is an HDF5 file and h5py is not installed")
        self.file = h5py.File(path, "r")
        return self.file


# ---- structure (io.py:13-50, 140-208) ------------------------------------------------------------------------------------------
def save_structure(fileobj, landmarks, *, group_name="structure", landmark_colors=None):
    landmarks = list(landmarks)
    with _Writer(fileobj, group_name) as g:
        views = sorted({obs.view for lm in landmarks for obs in lm.observations}, key=lambda v: v.frame_nr)
        observations = [obs for lm in landmarks for obs in lm.observations]
        view_to_index = {v: i for i, v in enumerate(views)}
        landmark_to_index = {lm: i for i, lm in enumerate(landmarks)}
        obs_to_index = {obs: i for i, obs in enumerate(observations)}
        gv = g.create_group("views")
        gv["frame_nr"] = np.array([v.frame_nr for v in views], dtype=np.int64)
        gv["t0"] = np.array([v.t0 for v in views], dtype=np.float64)
        gl = g.create_group("landmarks")
        gl["inverse_depth"] = np.array([lm.inverse_depth for lm in landmarks], dtype=np.float64)
        gl["ref_idx"] = np.array([obs_to_index[lm.reference] for lm in landmarks], dtype=np.int64)
        gl["color"] = np.vstack([landmark_colors[lm] for lm in landmarks]) if landmark_colors else np.empty((0, 3))
        go = g.create_group("observations")
        go["uv"] = np.vstack([obs.uv for obs in observations]) if observations else np.empty((0, 2))
        go["lm_idx"] = np.array([landmark_to_index[obs.landmark] for obs in observations], dtype=np.int64)
        go["v_idx"] = np.array([view_to_index[obs.view] for obs in observations], dtype=np.int64)


def load_structure(fileobj, group_name="structure"):
    """Returns (views, landmarks, landmark_colors)."""
    with _Reader(fileobj, group_name) as g:
        gv, gl, go = g["views"], g["landmarks"], g["observations"]
        views = [View(int(fnr), float(t0)) for fnr, t0 in zip(_value(gv["frame_nr"]), _value(gv["t0"]))]
        inverse_depth = _value(gl["inverse_depth"])
        landmarks = [Landmark() for _ in range(len(inverse_depth))]
        observations = [views[int(vi)].create_observation(landmarks[int(li)], uv)
                        for uv, li, vi in zip(_value(go["uv"]), _value(go["lm_idx"]), _value(go["v_idx"]))]
        for lm, invd, ref_idx in zip(landmarks, inverse_depth, _value(gl["ref_idx"])):
            lm.inverse_depth = float(invd)
            lm.reference = observations[int(ref_idx)]
        colors = _value(gl["color"])
        if len(colors) == len(landmarks) and len(landmarks) > 0:
            landmark_colors = {lm: c for lm, c in zip(landmarks, colors)}
        elif len(colors) == 0:
            landmark_colors = None
        else:
            raise IOError("Number of colors do not match!")
    return views, landmarks, landmark_colors


# ---- trajectories (io.py:52-103, 211-230) --------------------------------------------------------------------------------------
def _save_spline(group, spline):
    group["dt"] = spline.dt
    group["t0"] = spline.t0
    group["knots"] = np.stack([np.asarray(spline[i], float) for i in range(len(spline))]) if len(spline) else np.empty((0,))


def _load_spline(group, cls):
    instance = cls(float(_value(group["dt"])), float(_value(group["t0"])))
    for v in _value(group["knots"]):
        instance.append_knot(v)
    return instance


def save_trajectory(location, trajectory, group_name="trajectory"):
    with _Writer(location, group_name) as g:
        g["type"] = type(trajectory).__name__
        if isinstance(trajectory, SplitTrajectory):
            _save_spline(g.create_group("R3_spline"), trajectory.R3_spline)
            _save_spline(g.create_group("SO3_spline"), trajectory.SO3_spline)
        else:
            _save_spline(g, trajectory)


def load_trajectory(location, group_name="trajectory"):
    with _Reader(location, group_name) as g:
        name = _value(g["type"])
        if name == "SplitTrajectory":
            return SplitTrajectory(_load_spline(g["R3_spline"], UniformR3SplineTrajectory), _load_spline(g["SO3_spline"], UniformSO3SplineTrajectory))
        if name in _SPLINES:
            return _load_spline(g, _SPLINES[name])
        raise IOError(f"unknown trajectory type {name!r}")


def load_atan_camera(path):
    """io.py:106-115: camera calibration file with `size` (cols, rows), `readout`, `K`, `wc`, `lgamma`."""
    with _RootReader(path) as f:
        cols, rows = (int(x) for x in _value(f["size"]))
        return AtanCamera(rows, cols, float(_value(f["readout"])), _value(f["K"]), _value(f["wc"]), float(_value(f["lgamma"])))


# ---- object graph <-> flat index arrays -------------------------------------------------------------------
def flatten_structure(landmarks, include_reference=False):
    """The arrays of ktk_add_static_rs / ktk_add_newton_rs for every non-reference observation of `landmarks`, in landmark
    order then observation order (the order README.md:27-31 / python/tests/conftest.py:155-166 add measurements in):
    dict(obs_uv (n,2), obs_t0 (n,), ref_uv (n,2), ref_t0 (n,), lm_idx (n,) int32 into rho, rho (n_landmarks,),
    obs_view (n,) and ref_view (n_landmarks,) = frame numbers, locked (n_landmarks,) bool)."""
    landmarks = list(landmarks)
    rows = [(li, obs) for li, lm in enumerate(landmarks) for obs in lm.observations if include_reference or obs is not lm.reference]
    return dict(
        obs_uv=np.array([o.uv for _, o in rows], float).reshape(-1, 2), obs_t0=np.array([o.view.t0 for _, o in rows], float),
        ref_uv=np.array([landmarks[li].reference.uv for li, _ in rows], float).reshape(-1, 2),
        ref_t0=np.array([landmarks[li].reference.view.t0 for li, _ in rows], float),
        lm_idx=np.array([li for li, _ in rows], np.int32), rho=np.array([lm.inverse_depth for lm in landmarks], float),
        obs_view=np.array([o.view.frame_nr for _, o in rows], np.int64), ref_view=np.array([lm.reference.view.frame_nr for lm in landmarks], np.int64),
        locked=np.array([lm.locked for lm in landmarks], bool))


def structure_from_arrays(obs_uv, obs_t0, ref_uv, ref_t0, lm_idx, rho, obs_view=None, ref_view=None, locked=None):
    """Inverse of flatten_structure: (views, landmarks).  Views are identified by their frame number when obs_view / ref_view are
    given, otherwise by t0 (numbered in order of first appearance).  The reference observation of a landmark is created first."""
    obs_uv, ref_uv = np.asarray(obs_uv, float).reshape(-1, 2), np.asarray(ref_uv, float).reshape(-1, 2)
    obs_t0, ref_t0, lm_idx, rho = np.asarray(obs_t0, float), np.asarray(ref_t0, float), np.asarray(lm_idx), np.asarray(rho, float)
    views, landmarks = {}, [Landmark() for _ in range(len(rho))]

    def view(key, t0):
        if key not in views:
            views[key] = View(key if (obs_view is not None and ref_view is not None) else len(views), float(t0))
        return views[key]

    first = {}
    for i, li in enumerate(lm_idx):
        first.setdefault(int(li), i)
    for li, lm in enumerate(landmarks):
        lm.inverse_depth = float(rho[li])
        if locked is not None:
            lm.locked = bool(locked[li])
        if li in first:
            i = first[li]
            key = int(ref_view[li]) if (obs_view is not None and ref_view is not None) else float(ref_t0[i])
            lm.reference = view(key, ref_t0[i]).create_observation(lm, ref_uv[i])
    for i, li in enumerate(lm_idx):
        key = int(obs_view[i]) if (obs_view is not None and ref_view is not None) else float(obs_t0[i])
        view(key, obs_t0[i]).create_observation(landmarks[int(li)], obs_uv[i])
    return sorted(views.values(), key=lambda v: v.frame_nr), landmarks
