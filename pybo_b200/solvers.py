"""Inner-loop solvers: `solver(f, bounds, **kw) -> (xbest, fbest)`.

`solve_lbfgs` keeps the reference's contract (`pybo/solvers/lbfgs.py:17-68`):
score a candidate grid in one batched call, keep the `nbest` highest, refine each
with L-BFGS-B on -f using `f(x[None], grad=True)`, return `(x, f(x))`.
When `f` is a device-backed `ModelIndex` the grid is scored and reduced to its
top `nbest` on the GPU; any other callable takes the plain NumPy route.
"""

import numpy as np
import scipy.optimize

from .inits import init_uniform
from .utils import as_bounds

__all__ = ["solve_lbfgs"]


def _top_indices(values, k):
    """Indices of the k largest values, descending, lowest index first on ties."""
    order = np.lexsort((np.arange(len(values)), -values))
    return order[:k]


def solve_lbfgs(f, bounds, nbest=10, ngrid=10000, xgrid=None, rng=None, pick="first"):
    """Maximise `f` over the box.

    pick='first' reproduces the reference exactly: its final selection
    `result[np.argmin(generator)]` (lbfgs.py:65) always evaluates to `result[0]`,
    i.e. the refinement of the best grid point wins.  pick='best' returns the
    refined point with the highest value instead.
    """
    bounds = as_bounds(bounds)
    if xgrid is None:
        xgrid = init_uniform(bounds, ngrid, rng)
    else:
        xgrid = np.array(xgrid, dtype=float, ndmin=2)

    if getattr(f, "fused", False):
        starts, _ = f.best_of(xgrid, nbest)
    else:
        starts = _top_indices(np.asarray(f(xgrid, grad=False), dtype=float), nbest)

    def negated(x):
        fx, gx = f(x[None], grad=True)
        return -float(fx[0]), -np.asarray(gx[0], dtype=float)

    refined = []
    for x0 in xgrid[starts]:
        x, fmin, _ = scipy.optimize.fmin_l_bfgs_b(negated, x0, bounds=bounds)
        refined.append((x, fmin))

    if pick == "first":
        xbest, fmin = refined[0]
    elif pick == "best":
        xbest, fmin = min(refined, key=lambda r: r[1])
    else:
        raise ValueError("pick must be 'first' or 'best'")
    return xbest, -fmin
