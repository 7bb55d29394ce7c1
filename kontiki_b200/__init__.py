"""kontiki_b200 -- B200-native residual + Jacobian evaluation path of hovren/kontiki behind its Python surface.

The compute path is the CUDA library kontiki_b200/lib/libkontiki_b200.so (C ABI in include/kontiki_b200.h);
there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import Problem, KontikiError  # noqa: F401

__version__ = "0.1.0"

from . import io, measurements, sensors, sfm, trajectories, utils  # noqa: F401,E402
from .estimator import CallbackReturnType, IterationSummary, Summary, TerminationType, TrajectoryEstimator  # noqa: F401,E402
