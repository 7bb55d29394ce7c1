"""Structure-from-motion primitives with the reference's Python surface (python/src/kontiki/pysfm.cc:19-101,
cpplib/include/kontiki/sfm/{landmark,observation,view}{,_impl}.h).  The object graph stays on the host; the estimator
flattens it to index arrays for the device."""
import itertools

import numpy as np

_landmark_ids = itertools.count(0)       # process-global id counter starting at 0 (landmark.h:22-26)


class Landmark:
    def __init__(self):
        self.id = next(_landmark_ids)
        self.inverse_depth = 1.0
        self._reference = None
        self._observations = []
        self.locked = False              # landmark_impl.h:15

    @property
    def reference(self):
        if self._reference is None:
            raise RuntimeError("Landmark has no reference observation")
        return self._reference

    @reference.setter
    def reference(self, obs):
        if obs.landmark is not self:
            raise RuntimeError("Observation does not belong to this landmark")
        self._reference = obs

    @property
    def observations(self):
        return list(self._observations)

    def __repr__(self):
        return f"<Landmark num_obs={len(self._observations)}, inverse depth={self.inverse_depth}>"


class Observation:
    def __init__(self, view, landmark, uv):
        self._view, self._landmark = view, landmark
        self.uv = np.asarray(uv, float).copy()

    landmark = property(lambda self: self._landmark)
    view = property(lambda self: self._view)

    @property
    def is_reference(self):
        return self._landmark._reference is self

    def __repr__(self):
        return f"<Observation lm={self._landmark.id} f={self._view.frame_nr} t0={self._view.t0} uv={self.uv}>"


class View:
    def __init__(self, frame_nr, t0):
        self.frame_nr, self.t0 = int(frame_nr), float(t0)
        self._observations = []

    @property
    def observations(self):
        return list(self._observations)

    def create_observation(self, landmark, uv):              # view_impl.h:47-56: the only way to make an Observation
        obs = Observation(self, landmark, uv)
        self._observations.append(obs)
        landmark._observations.append(obs)
        return obs

    def remove_observation(self, obs):
        if obs not in self._observations:
            raise RuntimeError("Observation does not belong to this view")
        self._observations.remove(obs)
        obs.landmark._observations.remove(obs)
        if obs.landmark._reference is obs:
            obs.landmark._reference = None

    def __len__(self):
        return len(self._observations)

    def __repr__(self):
        return f"<View frame_nr={self.frame_nr} t0={self.t0}>"
