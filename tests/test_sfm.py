"""The reference's python/tests/test_sfm.py restated on kontiki_b200.sfm (the object graph the estimator flattens to
index arrays).  Pure host side: no GPU."""
import numpy as np
import pytest
from numpy.testing import assert_equal

from kontiki_b200.sfm import Landmark, View


def test_new_view():                                           # test_sfm.py:7-14
    v = View(34, 4.67)
    assert v.frame_nr == 34 and v.t0 == 4.67 and len(v) == 0 and len(v.observations) == 0


def test_view_add_observations():                              # :16-32
    lm1, lm2, v = Landmark(), Landmark(), View(0, 0.0)
    p1 = np.array([100, 200])
    v.create_observation(lm1, p1)
    assert len(v) == 1 and len(lm1.observations) == 1 and lm1.observations[0].view is v
    assert_equal(lm1.observations[0].uv, p1)
    assert len(lm2.observations) == 0
    v.create_observation(lm2, np.array([300, 499]))
    assert len(v) == 2 and len(lm2.observations) == 1


def test_remove_observations():                                # :34-48
    lm, v1, v2 = Landmark(), View(0, 0.0), View(1, 1.0)
    obs1 = v1.create_observation(lm, np.array([1, 2]))
    _ = v2.create_observation(lm, np.array([3, 4]))
    assert len(v1) == 1 and len(v2) == 1 and len(lm.observations) == 2
    v1.remove_observation(obs1)
    assert len(v1) == 0 and len(v2) == 1 and len(lm.observations) == 1


def test_remove_nonowned():                                    # :50-57
    lm, v, v_other = Landmark(), View(0, 0.0), View(1, 1.0)
    _ = v.create_observation(lm, np.array([1, 2]))
    obs_other = v_other.create_observation(lm, np.array([3, 4]))
    with pytest.raises(RuntimeError):
        v.remove_observation(obs_other)


def test_deleted_view_cleanup():                               # :59-69: a view owns its observations, landmarks only refer to them
    v = View(0, 0.0)
    landmarks = [Landmark() for _ in range(100)]
    for lm in landmarks:
        v.create_observation(lm, np.array([1, 1]))
        assert len(lm.observations) == 1
    del v
    for lm in landmarks:
        assert len(lm.observations) == 0


def test_new_landmark():                                       # :72-76
    lm = Landmark()
    assert len(lm.observations) == 0
    with pytest.raises(RuntimeError):
        lm.reference


def test_landmark_ids_unique():                                # :79-83
    assert len({lm.id for lm in [Landmark() for _ in range(1000)]}) == 1000


def test_landmark_reference_not_owned():                       # :86-97
    v, lm = View(0, 0.0), Landmark()
    obs_owned = v.create_observation(lm, np.array([1, 2]))
    obs_not_owned = v.create_observation(Landmark(), np.array([6, 7]))
    lm.reference = obs_owned
    assert lm.reference is obs_owned
    with pytest.raises(RuntimeError):
        lm.reference = obs_not_owned


def test_observation_is_reference():                           # :100-112
    views = [View(i, i) for i in range(4)]
    lm = Landmark()
    ref = views[0].create_observation(lm, np.array([1, 2]))
    lm.reference = ref
    not_refs = [v.create_observation(lm, np.array([1, 2])) for v in views]
    assert ref.is_reference and not any(obs.is_reference for obs in not_refs)


def test_remove_then_set_references():                         # :115-132
    rng = np.random.default_rng(0)
    landmarks = [Landmark() for _ in range(20)]
    views = [View(i, i) for i in range(30)]
    for v in views:
        for lm in landmarks:
            v.create_observation(lm, rng.uniform(0, 1000, size=2))
    for obs in [lm.observations[0] for lm in landmarks]:
        obs.view.remove_observation(obs)
    for lm in landmarks:
        lm.reference = lm.observations[0]
        assert lm.reference.view is views[1] and len(lm.observations) == 29


def test_dangling_view_raises():
    """observation_impl.h:22-28: an observation that outlives its view has no view any more."""
    lm, v = Landmark(), View(3, 0.1)
    obs = v.create_observation(lm, np.array([5.0, 6.0]))
    del v
    with pytest.raises(RuntimeError):
        obs.view
