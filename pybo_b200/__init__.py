"""pybo_b200 -- a B200-native GP Bayesian-optimisation inner loop behind pybo's
plugin surface (`solve_bayesopt`, `policies`, `solvers`, `recommenders`).
The numerics run in libbo_b200.so (hand-written sm_100a CUDA); see DESIGN.md."""

from .bayesopt import solve_bayesopt, init_model  # noqa: F401
from . import inits, models, policies, recommenders, solvers  # noqa: F401
from .models import make_gp, GP, MCMC  # noqa: F401

__all__ = ["solve_bayesopt", "init_model", "make_gp", "GP", "MCMC"]
