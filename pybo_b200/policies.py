"""Acquisition policies: `policy(model, bounds, X, **kw) -> index`, where
`index(X, grad=False)` scores an (M, d) batch and returns (M,) values, or
((M,), (M, d)) with `grad=True`.

Names, argument order and defaults follow the reference's
`pybo/policies/simple.py:13-74` (`EI(xi=0.0)`, `PI(xi=0.05)`,
`UCB(delta=0.1, xi=0.2)`, `Thompson(n=100, rng=None)`).  Each index is an object
rather than a bare closure so the solver can reach the fused device entry
points (`score + top-k` without copying M values to the host); calling it like
the reference's closure gives the reference's results.
"""

import numpy as np

from . import _lib

__all__ = ["EI", "PI", "UCB", "Thompson"]


class ModelIndex(object):
    """An acquisition bound to a private copy of the model (the reference takes
    `model.copy()` at simple.py:20,34,57 so later `add_data` calls cannot change
    an index that is still in use)."""

    def __init__(self, model, acq, param):
        self.model, self.acq, self.param = model, acq, float(param)

    # reference-style call --------------------------------------------------
    def __call__(self, X, grad=False):
        model = self.model
        if hasattr(model, "_acq"):                       # GPU-backed model: fused on device
            return model._acq(self.acq, self.param, X, grad)
        if self.acq == _lib.ACQ_EI:
            return model.get_improvement(self.param, X, grad)
        if self.acq == _lib.ACQ_PI:
            return model.get_tail(self.param, X, grad)
        post = model.predict(X, grad=grad)
        if self.acq == _lib.ACQ_MEAN:
            return (post[0], post[2]) if grad else post[0]
        mu, s2 = post[:2]                                # UCB, simple.py:64-72
        if not grad:
            return mu + np.sqrt(self.param * s2)
        dmu, ds2 = post[2:]
        return (mu + np.sqrt(self.param * s2),
                dmu + 0.5 * np.sqrt(self.param / s2[:, None]) * ds2)

    # fused device path -----------------------------------------------------
    @property
    def fused(self):
        return hasattr(self.model, "_ensure_fit") and self.model.ndata > 0

    def best_of(self, X, k):
        """Score X on the device and return the indices and values of the k best
        candidates (value descending, lowest index first among ties); only
        k values cross the PCIe bus."""
        ctx = self.model._ensure_fit()
        ctx.score(self.acq, self.param, X, want_values=False)
        return ctx.topk(min(int(k), len(X)))

    def best_of_sobol(self, bounds, M, k, start=0):
        """Same reduction over points [start, start + M) of the unscrambled Sobol sequence in `bounds`,
        generated on the device (no candidate array crosses the PCIe bus in either direction).  Returns
        (points (k, d), values (k,), sequence indices (k,))."""
        ctx = self.model._ensure_fit()
        bounds = np.array(bounds, dtype=np.float64, ndmin=2)
        d = bounds.shape[0]
        ctx.sobol(d, int(start), int(M), bounds, out="staged")
        ctx.score_staged(self.acq, self.param, int(M), want_best=False)
        idx, val = ctx.topk(min(int(k), int(M)))
        return _lib.sobol_points(d, idx + int(start), bounds), val, idx + int(start)


def _incumbent_target(model, X, xi):
    """max_i mu(x_i) + xi over the observed points (simple.py:21,35)."""
    return float(np.max(model.predict(X)[0])) + xi


def EI(model, _, X, xi=0.0):
    """Expected improvement over the best posterior mean at the data, plus xi."""
    model = model.copy()
    return ModelIndex(model, _lib.ACQ_EI, _incumbent_target(model, X, xi))


def PI(model, _, X, xi=0.05):
    """Probability of improving on the best posterior mean at the data by xi."""
    model = model.copy()
    return ModelIndex(model, _lib.ACQ_PI, _incumbent_target(model, X, xi))


def Thompson(model, _, __, n=100, rng=None):
    """One posterior function draw with n random features; its `.get` is the index."""
    return model.sample_f(n, rng).get


def ucb_beta(nobs, delta=0.1, xi=0.2):
    """beta = a + b log(d + 1) with the reference's constants (simple.py:58-66).
    The reference sets d = len(X), the number of observations, and so does this."""
    a = xi * 2 * np.log(np.pi ** 2 / 3 / delta)
    b = xi * (4 + nobs)
    return a + b * np.log(nobs + 1)


def UCB(model, _, X, delta=0.1, xi=0.2):
    """GP-UCB: mu + sqrt(beta s2)."""
    return ModelIndex(model.copy(), _lib.ACQ_UCB, ucb_beta(len(X), delta, xi))
