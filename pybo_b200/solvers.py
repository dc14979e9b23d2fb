"""Inner-loop solvers: `solver(f, bounds, **kw) -> (xbest, fbest)`.

`solve_lbfgs` keeps the reference's contract (`pybo/solvers/lbfgs.py:17-68`):
score a candidate grid in one batched call, keep the `nbest` highest, refine each
with L-BFGS-B on -f using `f(x[None], grad=True)`, return `(x, f(x))`.
When `f` is a device-backed `ModelIndex` the grid is scored and reduced to its
top `nbest` on the GPU; any other callable takes the plain NumPy route.

`solve_lbfgs_batched` (solver name 'lbfgs_batched') replaces the `nbest`
sequential SciPy runs (lbfgs.py:56-65: ~10 x 30 one-point callbacks, each streaming
the whole factor) by one projected L-BFGS that advances all starts in lockstep: every
iteration is ONE `f(X[nbest, d], grad=True)` call, served on the device by the
small-batch triangular matrix-vector path, which reads W once for all starts.
"""

import numpy as np
import scipy.optimize

from .inits import init_uniform
from .utils import as_bounds

__all__ = ["solve_lbfgs", "solve_lbfgs_batched", "batched_lbfgs"]


def _top_indices(values, k):
    """Indices of the k largest values, descending, lowest index first on ties."""
    order = np.lexsort((np.arange(len(values)), -values))
    return order[:k]


def _grid_starts(f, bounds, nbest, ngrid, xgrid, rng, grid):
    """The `nbest` best points of the candidate grid (reference lbfgs.py:42-51).  grid='uniform' is the
    reference's `init_uniform(bounds, ngrid, rng)`; grid='sobol' is the low-discrepancy grid its TODO
    (lbfgs.py:43-44) asks for -- for a device-backed index it is generated, scored and reduced on the
    GPU, so no candidate array is built on the host at all.  An explicit `xgrid` wins."""
    if grid not in ("uniform", "sobol"):
        raise ValueError("grid must be 'uniform' or 'sobol'")
    if xgrid is None and grid == "sobol":
        if getattr(f, "fused", False) and hasattr(f, "best_of_sobol"):
            return f.best_of_sobol(bounds, ngrid, nbest)[0]
        from ._lib import sobol_points
        xgrid = sobol_points(len(bounds), np.arange(ngrid), bounds)
    elif xgrid is None:
        xgrid = init_uniform(bounds, ngrid, rng)
    else:
        xgrid = np.array(xgrid, dtype=float, ndmin=2)
    if getattr(f, "fused", False):
        starts, _ = f.best_of(xgrid, nbest)
    else:
        starts = _top_indices(np.asarray(f(xgrid, grad=False), dtype=float), nbest)
    return xgrid[starts]


def solve_lbfgs(f, bounds, nbest=10, ngrid=10000, xgrid=None, rng=None, pick="first", grid="uniform"):
    """Maximise `f` over the box.

    pick='first' reproduces the reference exactly: its final selection
    `result[np.argmin(generator)]` (lbfgs.py:65) always evaluates to `result[0]`,
    i.e. the refinement of the best grid point wins.  pick='best' returns the
    refined point with the highest value instead.
    """
    bounds = as_bounds(bounds)
    x0 = _grid_starts(f, bounds, nbest, ngrid, xgrid, rng, grid)

    def negated(x):
        fx, gx = f(x[None], grad=True)
        return -float(fx[0]), -np.asarray(gx[0], dtype=float)

    refined = []
    for start in x0:
        x, fmin, _ = scipy.optimize.fmin_l_bfgs_b(negated, start, bounds=bounds)
        refined.append((x, fmin))

    if pick == "first":
        xbest, fmin = refined[0]
    elif pick == "best":
        xbest, fmin = min(refined, key=lambda r: r[1])
    else:
        raise ValueError("pick must be 'first' or 'best'")
    return xbest, -fmin


def batched_lbfgs(f, x0, bounds, maxiter=100, history=10, pgtol=1e-5, ftol=2.2e-9, c1=1e-4, maxls=20):
    """Maximise `f` from the B starting points `x0` (B, d) inside the box, all starts in lockstep.

    Projected limited-memory BFGS on -f: the two-loop recursion runs per start on the variables
    that are free at the current point (not pinned at a bound by the gradient), the step is
    projected back onto the box, and an Armijo backtracking search accepts it; every function /
    gradient evaluation is a single batched call `f(X, grad=True) -> ((B,), (B, d))`.
    Stopping per start mirrors `fmin_l_bfgs_b`'s defaults: projected-gradient sup-norm <= pgtol
    or relative decrease <= ftol (factr = 1e7).  Returns (x (B, d), f(x) (B,), evaluations).
    """
    bounds = as_bounds(bounds)
    lo, hi = bounds[:, 0], bounds[:, 1]
    x = np.clip(np.array(x0, dtype=float, ndmin=2), lo, hi)
    B, d = x.shape

    def evaluate(points):
        fx, gx = f(points, grad=True)
        return -np.asarray(fx, dtype=float).reshape(B), -np.asarray(gx, dtype=float).reshape(B, d)

    fx, gx = evaluate(x)
    nev = 1
    S = np.zeros((history, B, d))
    Y = np.zeros((history, B, d))
    rho = np.zeros((history, B))                       # 1 / (s.y); 0 marks an unused / skipped pair
    active = np.ones(B, dtype=bool)
    step0 = np.ones(B)

    def projected_gradient(x, g):
        pg = g.copy()
        pg[(x <= lo) & (g > 0)] = 0.0
        pg[(x >= hi) & (g < 0)] = 0.0
        return pg

    for it in range(maxiter):
        pg = projected_gradient(x, gx)
        active &= np.max(np.abs(pg), axis=1) > pgtol
        if not active.any():
            break
        free = pg != 0.0
        # two-loop recursion on the free variables, newest pair first
        q = np.where(free, gx, 0.0)
        order = [(it - 1 - k) % history for k in range(min(it, history))]
        a = np.zeros((len(order), B))
        for k, h in enumerate(order):
            a[k] = rho[h] * np.sum(S[h] * q, axis=1)
            q -= a[k][:, None] * Y[h]
        if order:
            h = order[0]
            yy = np.sum(Y[h] * Y[h], axis=1)
            gamma = np.where((rho[h] > 0) & (yy > 0), 1.0 / np.maximum(rho[h] * yy, 1e-300), 1.0)
        else:
            gamma = 1.0 / np.maximum(np.max(np.abs(pg), axis=1), 1e-12) * np.minimum(1.0, np.max(hi - lo))
        r = q * np.asarray(gamma)[:, None] if np.ndim(gamma) else q * gamma
        for k, h in reversed(list(enumerate(order))):
            b = rho[h] * np.sum(Y[h] * r, axis=1)
            r += (a[k] - b)[:, None] * S[h]
        dirn = np.where(free, -r, 0.0)
        slope = np.sum(dirn * gx, axis=1)
        bad = ~(slope < 0)                              # not a descent direction: steepest descent on the free set
        if bad.any():
            dirn[bad] = -pg[bad]
            slope[bad] = np.sum(dirn[bad] * gx[bad], axis=1)
        # Armijo backtracking along the projected path, all starts at once
        t = np.where(it == 0, np.minimum(1.0, 1.0 / np.maximum(np.linalg.norm(dirn, axis=1), 1e-300)), step0)
        searching = active.copy()
        xn, fn, gn = x.copy(), fx.copy(), gx.copy()
        for _ in range(maxls):
            trial = np.clip(x + t[:, None] * dirn, lo, hi)
            trial[~searching] = xn[~searching]
            ft, gt = evaluate(trial)
            nev += 1
            decrease = c1 * np.sum(gx * (trial - x), axis=1)
            ok = searching & np.isfinite(ft) & (ft <= fx + decrease)
            xn[ok], fn[ok], gn[ok] = trial[ok], ft[ok], gt[ok]
            searching &= ~ok
            if not searching.any():
                break
            t[searching] *= 0.5
        active &= ~searching                            # line search failed: stop this start where it is
        s_new, y_new = xn - x, gn - gx
        sy = np.sum(s_new * y_new, axis=1)
        good = active & (sy > 1e-10 * np.sum(y_new * y_new, axis=1)) & (sy > 0)
        h = it % history
        S[h], Y[h] = np.where(good[:, None], s_new, 0.0), np.where(good[:, None], y_new, 0.0)
        rho[h] = np.where(good, 1.0 / np.where(good, sy, 1.0), 0.0)
        small = (fx - fn) <= ftol * np.maximum(np.maximum(np.abs(fx), np.abs(fn)), 1.0)
        x, fx, gx = xn, fn, gn
        active &= ~small
    return x, -fx, nev


def solve_lbfgs_batched(f, bounds, nbest=10, ngrid=10000, xgrid=None, rng=None, pick="best", maxiter=100,
                        grid="uniform"):
    """Maximise `f` over the box: grid scoring and top-`nbest` selection as in `solve_lbfgs`
    (reference lbfgs.py:42-51), then all `nbest` starts refined together by `batched_lbfgs`.
    pick='best' (default) returns the highest refined value; pick='first' keeps the reference's
    `result[0]` selection (lbfgs.py:65).  Same return contract: `(xbest (d,), f(xbest))`."""
    bounds = as_bounds(bounds)
    x, fx, _ = batched_lbfgs(f, _grid_starts(f, bounds, nbest, ngrid, xgrid, rng, grid), bounds, maxiter=maxiter)
    if pick == "first":
        k = 0
    elif pick == "best":
        k = int(np.argmax(fx))
    else:
        raise ValueError("pick must be 'first' or 'best'")
    return x[k], float(fx[k])
