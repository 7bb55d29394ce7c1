"""Structure-from-motion primitives with the reference's Python surface (python/src/kontiki/pysfm.cc:19-101,
cpplib/include/kontiki/sfm/{landmark,observation,view}{,_impl}.h).  The object graph stays on the host; the estimator
flattens it to index arrays for the device."""
import itertools
import weakref

import numpy as np

_landmark_ids = itertools.count(0)       # process-global id counter starting at 0 (landmark.h:22-26)


class Landmark:
    def __init__(self):
        self.id = next(_landmark_ids)
        self.inverse_depth = 1.0
        self._reference = None
        self._observations = []
        self.locked = False              # landmark_impl.h:15

    # Ownership as in the reference: a View owns its observations (shared_ptr, view.h:30-33), a Landmark only refers to them
    # (weak_ptr, landmark.h:46-49), so observations of a deleted view drop out of `observations` (landmark_impl.h:40-52).
    @property
    def reference(self):
        ref = self._reference() if self._reference is not None else None
        if ref is None:
            raise RuntimeError("Landmark has no reference observation")
        return ref

    @reference.setter
    def reference(self, obs):
        if obs.landmark is not self:
            raise RuntimeError("Observation does not belong to this landmark")
        self._reference = weakref.ref(obs)

    @property
    def observations(self):
        alive = [o for o in (w() for w in self._observations) if o is not None]
        self._observations = [weakref.ref(o) for o in alive]
        return alive

    def __repr__(self):
        return f"<Landmark num_obs={len(self.observations)}, inverse depth={self.inverse_depth}>"


class Observation:
    def __init__(self, view, landmark, uv):
        self._view, self._landmark = weakref.ref(view), landmark      # observation.h:28-31: weak_ptr to the view, shared_ptr to the landmark
        self.uv = np.asarray(uv, float).copy()

    landmark = property(lambda self: self._landmark)

    @property
    def view(self):
        v = self._view()
        if v is None:
            raise RuntimeError("Observation's view no longer exists")      # observation_impl.h:22-28
        return v

    @property
    def is_reference(self):
        ref = self._landmark._reference
        return ref is not None and ref() is self

    def __repr__(self):
        return f"<Observation lm={self._landmark.id} f={self.view.frame_nr} t0={self.view.t0} uv={self.uv}>"


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
        landmark._observations.append(weakref.ref(obs))
        return obs

    def remove_observation(self, obs):
        if not any(o is obs for o in self._observations):
            raise RuntimeError("Observation does not belong to this view")      # view_impl.h:58-70
        self._observations = [o for o in self._observations if o is not obs]
        lm = obs.landmark
        lm._observations = [w for w in lm._observations if w() is not None and w() is not obs]
        if lm._reference is not None and lm._reference() is obs:
            lm._reference = None

    def __len__(self):
        return len(self._observations)

    def __repr__(self):
        return f"<View frame_nr={self.frame_nr} t0={self.t0}>"
