"""The Bayesian-optimisation meta solver (host glue, Python 3).

Keeps the public behaviour of the reference's `pybo/bayesopt.py`:
`solve_bayesopt(...) -> (xbest, model, Info)` with components given as a
callable, a lowercase name or a `(name_or_callable, kwargs)` pair
(bayesopt.py:125-176, 193-287), `init_model` with the reference's default
hyper-parameters and priors (bayesopt.py:60-120), and a pickle checkpoint after
every objective evaluation (bayesopt.py:39-55).  The model it builds is the
GPU-backed one from `pybo_b200.models`.
"""

import collections
import functools
import inspect
import os
import pickle

import numpy as np

from . import inits
from . import models
from . import policies
from . import recommenders
from . import solvers
from .utils import rstate, as_bounds

__all__ = ["solve_bayesopt", "init_model"]

Info = collections.namedtuple("Info", ["x", "y", "xbest"])


# ---- checkpoint helpers -----------------------------------------------------

def safe_dump(model, info, filename=None):
    """Pickle `(model, info)` to `filename` (no-op when it is None).  Written to a
    temporary file first so an interrupted run never leaves a truncated log."""
    if filename is None:
        return
    tmp = filename + ".tmp"
    with open(tmp, "wb") as fp:
        pickle.dump((model, info), fp)
    os.replace(tmp, filename)


def safe_load(filename=None):
    """Load a checkpoint, or `(None, Info([], [], []))` when there is none."""
    if filename is not None and os.path.exists(filename):
        with open(filename, "rb") as fp:
            return pickle.load(fp)
    return None, Info([], [], [])


# ---- model bootstrap --------------------------------------------------------

def init_model(f, bounds, ninit=None, design="latin", log=None, rng=None, kernel="se"):
    """Evaluate an initial design and build the default model: a GP with
    sn2 = 1e-6, rho = range(y) (1 if < 0.1), ell = width / 4, bias = mean(y),
    the reference's hyper-priors, wrapped in `MCMC(n=10, burn=100)`.
    Resumes from `log`: only design points whose y is still NaN are evaluated."""
    rng = rstate(rng)
    bounds = as_bounds(bounds)
    model, info = safe_load(log)
    if model is not None:
        return model
    if len(info.x) == 0:
        ninit = 3 * len(bounds) if ninit is None else ninit
        make_design = getattr(inits, "init_" + design)
        info.x.extend(make_design(bounds, ninit, rng))
        info.y.extend([np.nan] * ninit)
    for i, x in enumerate(info.x):
        if np.isnan(info.y[i]):
            info.y[i] = f(x)
        safe_dump(None, info, filename=log)

    ys = np.asarray(info.y, dtype=float)
    sn2 = 1e-6
    rho = float(ys.max() - ys.min()) if len(ys) > 1 else 1.0
    if rho < 1e-1:
        rho = 1.0
    ell = 0.25 * (bounds[:, 1] - bounds[:, 0])
    bias = float(ys.mean()) if len(ys) else 0.0

    model = models.make_gp(sn2, rho, ell, bias, kernel=kernel)
    model.params["like.sn2"].set_prior("horseshoe", 0.1)
    model.params["kern.rho"].set_prior("lognormal", np.log(rho), 1.0)
    model.params["kern.ell"].set_prior("uniform", ell / 100, ell * 10)
    model.params["mean.bias"].set_prior("normal", bias, rho)
    model.add_data(info.x, info.y)
    model = models.MCMC(model, n=10, burn=100, rng=rng)
    safe_dump(model, info, filename=log)
    return model


# ---- plugin resolution ------------------------------------------------------

def get_component(value, module, rng, lstrip=""):
    """Resolve a component spec to a callable with its kwargs (and the shared rng,
    if it takes one) bound.  Unknown names or kwargs raise ValueError."""
    kwargs = {}
    if isinstance(value, (list, tuple)):
        try:
            value, kwargs = value
            kwargs = dict(kwargs)
        except (ValueError, TypeError):
            raise ValueError("invalid component: {!r}".format(value))

    if callable(value):
        func = value
    else:
        table = {}
        for fname in module.__all__:
            short = fname[len(lstrip):] if fname.startswith(lstrip) else fname
            table[short.lower()] = getattr(module, fname)
        if value not in table:
            raise ValueError("invalid component: {!s}".format(value))
        func = table[value]

    params = inspect.signature(func).parameters
    optional = {k for k, p in params.items() if p.default is not inspect.Parameter.empty}
    optional.discard("rng")
    unknown = set(kwargs) - optional
    if unknown:
        raise ValueError("unknown arguments for {}: {}".format(
            getattr(func, "__name__", repr(func)), ", ".join(sorted(unknown))))
    if "rng" in params:
        kwargs["rng"] = rng
    return functools.partial(func, **kwargs) if kwargs else func


# ---- progress formatting ----------------------------------------------------

def _fmt_array(a):
    return np.array2string(np.asarray(a), formatter=dict(float="{: .3f}".format, int="{:03d}".format))


# ---- the loop ---------------------------------------------------------------

def solve_bayesopt(objective, bounds, model=None, niter=100, policy="ei", solver="lbfgs",
                   recommender="latent", ninit=None, verbose=False, log=None, rng=None):
    """Maximise `objective` over the box `bounds` by GP Bayesian optimisation.

    Returns `(xbest, model, info)` with `info = Info(x, y, xbest)` arrays holding
    the queries, observations and per-iteration recommendations.  Iteration
    order is the reference's (bayesopt.py:262-276): policy -> solver -> objective
    -> `model.add_data` -> recommender (called with the query list *before* the
    newest point is appended) -> checkpoint.
    """
    rng = rstate(rng)
    bounds = as_bounds(bounds)
    policy = get_component(policy, policies, rng)
    solver = get_component(solver, solvers, rng, lstrip="solve_")
    recommender = get_component(recommender, recommenders, rng, lstrip="best_")

    saved, info = safe_load(log)
    if model is None and saved is None:
        # as in the reference (bayesopt.py:243-246) `info` was loaded *before* the
        # model is initialised, so the design points live in the model only and the
        # trace starts with the mid-point query below.
        model = init_model(objective, bounds, ninit, log=log, rng=rng)
    else:
        model = saved if saved is not None else model.copy()

    if len(info.x) == 0:
        x = inits.init_middle(bounds)[0]
        y = objective(x)
        info.x.append(x)
        info.y.append(y)
        model.add_data(x, y)
        safe_dump(model, info, filename=log)

    xbest = info.xbest[-1] if len(info.xbest) else None
    for i in range(len(info.xbest), niter):
        index = policy(model, bounds, info.x)
        x, _ = solver(index, bounds)
        del index                       # releases the policy's private model copy: add_data can then
        y = objective(x)                # append to the device factors in place (bo_append) instead of refitting
        model.add_data(x, y)
        xbest = recommender(model, bounds, info.x)
        info.x.append(x)
        info.y.append(y)
        info.xbest.append(xbest)
        safe_dump(model, info, filename=log)
        if verbose:
            print("i={:03d}, x={}, y={: .3f}, xbest={}".format(i, _fmt_array(x), float(y), _fmt_array(xbest)))

    info = Info(*[np.array(v) for v in info])
    return xbest, model, info
