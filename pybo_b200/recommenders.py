"""Recommenders: `recommender(model, bounds, X) -> xbest`
(reference `pybo/recommenders.py:14-35`)."""

import numpy as np

from . import _lib
from . import solvers
from .policies import ModelIndex

__all__ = ["best_latent", "best_incumbent"]


def best_latent(model, bounds, X):
    """Maximiser of the posterior mean, found by `solve_lbfgs` seeded with the
    observed points as its grid (recommenders.py:19-26)."""
    xbest, _ = solvers.solve_lbfgs(ModelIndex(model, _lib.ACQ_MEAN, 0.0), bounds, xgrid=X)
    return xbest


def best_incumbent(model, _, X):
    """The observed point with the highest posterior mean (first one on ties)."""
    mu, _ = model.predict(X)
    return np.asarray(X)[int(np.argmax(mu))]
