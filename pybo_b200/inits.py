"""Initial designs: `init_*(bounds, n, rng) -> (n, d)`.

Same names, defaults (n = 3 d) and RNG call order as the reference's
`pybo/inits/methods.py:17-77`, so a shared seed gives the same design:
uniform draws `rng.rand(n, d)`; Latin hypercube draws `rng.rand(n, d)` and then
one `rng.permutation` per column.  `init_sobol` uses SciPy's unscrambled Sobol
generator (the reference vendors its own direction-number table,
inits/sobol.py, whose points differ from SciPy's beyond the second dimension).
"""

import numpy as np
from scipy.stats import qmc

from .utils import rstate, as_bounds

__all__ = ["init_middle", "init_uniform", "init_latin", "init_sobol"]


def _unpack(bounds, n):
    b = as_bounds(bounds)
    lo, width = b[:, 0], b[:, 1] - b[:, 0]
    return lo, width, (3 * len(b) if n is None else int(n))


def init_middle(bounds):
    """The centre of the box as a (1, d) design."""
    b = as_bounds(bounds)
    return (0.5 * (b[:, 0] + b[:, 1]))[None, :]


def init_uniform(bounds, n=None, rng=None):
    lo, width, n = _unpack(bounds, n)
    return lo + width * rstate(rng).rand(n, len(lo))


def init_latin(bounds, n=None, rng=None):
    rng = rstate(rng)
    lo, width, n = _unpack(bounds, n)
    X = lo + width * (np.arange(n)[:, None] + rng.rand(n, len(lo))) / n
    for k in range(len(lo)):
        X[:, k] = rng.permutation(X[:, k])
    return X


def init_sobol(bounds, n=None, rng=None):
    rng = rstate(rng)
    lo, width, n = _unpack(bounds, n)
    skip = rng.randint(100, 200)
    engine = qmc.Sobol(d=len(lo), scramble=False)
    if skip:
        engine.fast_forward(skip)
    return lo + width * engine.random(n)
