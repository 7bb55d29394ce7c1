"""Small helpers of the reference's Python package that its own tests and user scripts lean on (python/kontiki/utils.py:3-61):
a time, or a time span, at which a trajectory is certainly valid.  Pure host side."""
import math


def _bounds(trajectory):
    tmin, tmax = trajectory.valid_time
    if not tmax > tmin:
        raise ValueError(f"No safe time: the valid time interval is empty ({tmin}, {tmax})")
    return tmin, tmax


def safe_time(trajectory):
    """A time instance inside the trajectory's valid time: the middle of a finite interval, one unit inside a half-open one."""
    tmin, tmax = _bounds(trajectory)
    lo, hi = math.isfinite(tmin), math.isfinite(tmax)
    t = 0.5 * (tmin + tmax) if (lo and hi) else (tmin + 1.0 if lo else (tmax - 1.0 if hi else 42.0))
    if not math.isfinite(t):
        raise ValueError("No safe time: result was not finite")
    return t


def safe_time_span(trajectory, length, *, allow_shorter=False):
    """A span (t1, t2) of the given length inside the trajectory's valid time (the whole valid time if allow_shorter and it is shorter)."""
    tmin, tmax = _bounds(trajectory)
    lo, hi = math.isfinite(tmin), math.isfinite(tmax)
    if lo and hi:
        if tmax - tmin < length:
            if not allow_shorter:
                raise ValueError("No safe time span: trajectory is too short")
            span = (tmin, tmax)
        else:
            span = (tmin, tmin + length)
    elif lo:
        span = (tmin, tmin + length)
    elif hi:
        span = (tmax - length, tmax)
    else:
        span = (42.0, 42.0 + length)
    if not all(math.isfinite(v) for v in span):
        raise ValueError("No safe time span: got non-finite result")
    return span
