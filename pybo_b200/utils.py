"""Small host-side helpers shared by the plugin layer."""

import numpy as np

__all__ = ["rstate", "as_bounds"]


def rstate(rng=None):
    """Seed-or-state -> `numpy.random.RandomState` (same contract as the
    reference's `pybo.utils.rstate`, utils.py:16-24: an existing RandomState is
    passed through untouched so one stream threads through every component)."""
    return rng if isinstance(rng, np.random.RandomState) else np.random.RandomState(rng)


def as_bounds(bounds):
    """(d, 2) float array of (lower, upper) per dimension."""
    b = np.array(bounds, dtype=float, ndmin=2)
    if b.ndim != 2 or b.shape[1] != 2:
        raise ValueError("bounds must have shape (d, 2)")
    return b
