"""Host-side logic that needs no GPU: plugin resolution, designs, the solver on a
plain callable, the BO loop driven by the oracle model (BASELINE config 1), the
C-ABI export list, and that the product refuses to run without CUDA."""

import os
import pickle
import re

import numpy as np
import scipy.optimize
import pytest

import pybo_b200
from pybo_b200 import _lib, bayesopt, inits, models, policies, recommenders, solvers
from oracle import GPOracle, ucb_beta, ucb_index

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI ----------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "bo_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(bo_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/bo_b200.h must compile as C99 (no C++ or torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "%s"\nint main(void) { bo_ctx *c = 0; (void)c; return BO_OK; }\n'
                   % os.path.join(ROOT, "include", "bo_b200.h"))
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-c", str(src), "-o",
                          str(tmp_path / "hdr.o")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BackendError):
        _lib.Context(0)
    gp = models.make_gp(1e-6, 1.0, [0.25, 0.25], 0.0)
    gp.add_data(np.random.rand(5, 2), np.random.rand(5))
    with pytest.raises(_lib.BackendError):
        gp.predict(np.random.rand(3, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pybo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f


# ---- plugin resolution (reference bayesopt.py:125-176) -----------------------
def test_get_component_by_name_callable_and_tuple():
    rng = np.random.RandomState(0)
    assert bayesopt.get_component("ei", policies, rng) is policies.EI
    assert bayesopt.get_component("thompson", policies, rng).keywords == {"rng": rng}
    f = bayesopt.get_component(("ucb", {"delta": 0.5}), policies, rng)
    assert f.func is policies.UCB and f.keywords == {"delta": 0.5}
    s = bayesopt.get_component("lbfgs", solvers, rng, lstrip="solve_")
    assert s.func is solvers.solve_lbfgs and s.keywords["rng"] is rng
    assert bayesopt.get_component("incumbent", recommenders, rng, lstrip="best_") is recommenders.best_incumbent
    g = lambda model, bounds, X: None
    assert bayesopt.get_component(g, policies, rng) is g


def test_get_component_errors():
    rng = np.random.RandomState(0)
    with pytest.raises(ValueError):
        bayesopt.get_component("nope", policies, rng)
    with pytest.raises(ValueError):
        bayesopt.get_component(("ei", {"bogus": 1}), policies, rng)
    with pytest.raises(ValueError):
        bayesopt.get_component(("ei", {"rng": 1}), policies, rng)      # rng is not user-settable
    with pytest.raises(ValueError):
        bayesopt.get_component(("ei", 1, 2), policies, rng)


# ---- designs (reference inits/methods.py:17-77) -------------------------------
def test_designs_follow_reference_rng_order():
    b = np.array([[-5, 10.0], [0, 15.0], [1, 2.0]])
    lo, w = b[:, 0], b[:, 1] - b[:, 0]
    assert np.allclose(inits.init_middle(b), [[2.5, 7.5, 1.5]])
    r = np.random.RandomState(3)
    assert np.array_equal(inits.init_uniform(b, 7, rng=3), lo + w * r.rand(7, 3))
    assert inits.init_uniform(b).shape == (9, 3)
    r = np.random.RandomState(4)
    X = lo + w * (np.arange(6)[:, None] + r.rand(6, 3)) / 6
    for k in range(3):
        X[:, k] = r.permutation(X[:, k])
    assert np.array_equal(inits.init_latin(b, 6, rng=4), X)
    S = inits.init_sobol(b, 16, rng=0)
    assert S.shape == (16, 3) and np.all(S >= lo) and np.all(S <= b[:, 1])
    # each Latin column hits every stratum exactly once
    L = (inits.init_latin(b, 10, rng=1) - lo) / w
    assert all(sorted(np.floor(L[:, k] * 10).astype(int)) == list(range(10)) for k in range(3))


# ---- solver on a plain callable ---------------------------------------------
def test_solve_lbfgs_plain_callable_and_quirk():
    centre = np.array([0.3, -0.2])

    def f(X, grad=False):
        X = np.array(X, ndmin=2)
        val = -np.sum((X - centre) ** 2, axis=1)
        return (val, -2 * (X - centre)) if grad else val

    b = np.array([[-1, 1.0], [-1, 1.0]])
    x, fx = solvers.solve_lbfgs(f, b, ngrid=500, rng=0)
    assert np.allclose(x, centre, atol=1e-5) and abs(fx) < 1e-9
    # multimodal: pick='first' refines the best grid point only (reference quirk)
    def g(X, grad=False):
        X = np.array(X, ndmin=2)
        a = np.exp(-50 * np.sum((X - 0.5) ** 2, axis=1))
        c = 2 * np.exp(-400 * np.sum((X + 0.5) ** 2, axis=1))
        val = a + c
        if not grad:
            return val
        return val, (-100 * (X - 0.5) * a[:, None] - 800 * (X + 0.5) * c[:, None])
    grid = np.array([[0.45, 0.5], [-0.4, -0.4], [0.0, 0.0]])
    x1, f1 = solvers.solve_lbfgs(g, b, xgrid=grid, nbest=3, pick="first")
    x2, f2 = solvers.solve_lbfgs(g, b, xgrid=grid, nbest=3, pick="best")
    assert np.allclose(x1, 0.5, atol=1e-4) and abs(f1 - 1.0) < 1e-6
    assert f2 >= f1


def test_batched_multistart_solver_matches_sequential_lbfgs():
    """solve_lbfgs_batched: all starts advance in lockstep, one batched f call per step;
    same optima as the reference's sequential fmin_l_bfgs_b runs, in far fewer callbacks."""
    calls = {"n": 0, "rows": 0}

    def f(X, grad=False):
        X = np.array(X, ndmin=2)
        calls["n"] += 1
        calls["rows"] += len(X)
        v = np.sin(3 * X[:, 0]) * np.cos(2 * X[:, 1]) - 0.1 * np.sum((X - 0.3) ** 2, axis=1)
        if not grad:
            return v
        g = np.empty_like(X)
        g[:, 0] = 3 * np.cos(3 * X[:, 0]) * np.cos(2 * X[:, 1]) - 0.2 * (X[:, 0] - 0.3)
        g[:, 1] = -2 * np.sin(3 * X[:, 0]) * np.sin(2 * X[:, 1]) - 0.2 * (X[:, 1] - 0.3)
        return v, g

    b = np.array([[-2, 2.0], [-1, 3.0]])
    grid = b[:, 0] + (b[:, 1] - b[:, 0]) * np.random.RandomState(0).rand(2000, 2)
    xa, fa = solvers.solve_lbfgs(f, b, xgrid=grid, pick="best")
    seq_calls = calls["n"]
    calls["n"] = 0
    xb, fb = solvers.solve_lbfgs_batched(f, b, xgrid=grid)
    assert np.allclose(xa, xb, atol=1e-5) and abs(fa - fb) < 1e-9
    assert calls["n"] < seq_calls / 3
    # every start climbs to a stationary point of the box-constrained problem (which basin a start
    # ends in may differ from SciPy's: both are local searches with different step rules)
    starts = grid[np.argsort(-f(grid))[::200][:7]]
    x, fx, nev = solvers.batched_lbfgs(f, starts, b)
    v, g = f(x, grad=True)
    pg = np.where(((x <= b[:, 0]) & (g < 0)) | ((x >= b[:, 1]) & (g > 0)), 0.0, g)
    assert np.allclose(v, fx) and np.all(fx >= f(starts) - 1e-12) and np.max(np.abs(pg)) < 1e-4
    assert np.all(x >= b[:, 0]) and np.all(x <= b[:, 1])
    # optimum on the boundary of the box; pick='first' keeps the reference's result[0] choice
    def g(X, grad=False):
        X = np.array(X, ndmin=2)
        v = X[:, 0] + 0.5 * X[:, 1] - X[:, 1] ** 2
        return (v, np.column_stack([np.ones(len(X)), 0.5 - 2 * X[:, 1]])) if grad else v
    x, fx = solvers.solve_lbfgs_batched(g, [[0, 1], [0, 1]], ngrid=100, rng=0)
    assert np.allclose(x, [1.0, 0.25], atol=1e-6) and abs(fx - 1.0625) < 1e-10
    x1, f1 = solvers.solve_lbfgs_batched(g, [[0, 1], [0, 1]], ngrid=100, rng=0, pick="first")
    assert abs(f1 - 1.0625) < 1e-10
    with pytest.raises(ValueError):
        solvers.solve_lbfgs_batched(g, [[0, 1], [0, 1]], ngrid=10, rng=0, pick="nope")
    s = bayesopt.get_component("lbfgs_batched", solvers, np.random.RandomState(0), lstrip="solve_")
    assert s.func is solvers.solve_lbfgs_batched


def test_sobol_grid_host_restatement_and_solver_option():
    """`grid='sobol'` (the low-discrepancy grid of the TODO at reference lbfgs.py:43-44): the host restatement
    of the device generator is bit-identical to SciPy's unscrambled sequence, scaled to the box."""
    from scipy.stats import qmc
    for d in (1, 3, 16):
        ref = qmc.Sobol(d=d, scramble=False).random_base2(9)
        assert np.array_equal(_lib.sobol_points(d, np.arange(512)), ref)
        assert np.array_equal(_lib.sobol_points(d, [7, 300, 511]), ref[[7, 300, 511]])
    b = np.array([[-5, 10.0], [0, 15]])
    pts = _lib.sobol_points(2, np.arange(64), b)
    assert np.all(pts >= b[:, 0]) and np.all(pts <= b[:, 1])
    assert np.allclose(pts, b[:, 0] + 15.0 * qmc.Sobol(d=2, scramble=False).random_base2(6))

    def f(X, grad=False):
        X = np.array(X, ndmin=2)
        v = -np.sum((X - np.array([2.0, 3.0])) ** 2, axis=1)
        return (v, -2 * (X - np.array([2.0, 3.0]))) if grad else v
    for solve in (solvers.solve_lbfgs, solvers.solve_lbfgs_batched):
        x, fx = solve(f, b, ngrid=256, grid="sobol")
        assert np.allclose(x, [2.0, 3.0], atol=1e-5) and abs(fx) < 1e-9
        with pytest.raises(ValueError):
            solve(f, b, ngrid=16, grid="halton")


def test_policies_match_reference_formulas_on_oracle_model():
    rng = np.random.RandomState(0)
    gp = GPOracle(1e-4, 1.3, [0.3, 0.4], 0.1, "se")
    X = rng.rand(15, 2)
    gp.add_data(X, np.sin(X.sum(1)))
    Xc = rng.rand(20, 2)
    ei = policies.EI(gp, None, list(X), xi=0.01)
    t = gp.predict(X)[0].max() + 0.01
    assert np.array_equal(ei(Xc), gp.get_improvement(t, Xc))
    v, g = ei(Xc, grad=True)
    assert g.shape == (20, 2)
    pi = policies.PI(gp, None, list(X))
    assert np.array_equal(pi(Xc), gp.get_tail(gp.predict(X)[0].max() + 0.05, Xc))
    ucb = policies.UCB(gp, None, list(X))
    assert abs(ucb.param - ucb_beta(15)) < 1e-14
    mu, s2, dmu, ds2 = gp.predict(Xc, grad=True)
    u, du = ucb(Xc, grad=True)
    ru, rdu = ucb_index(ucb_beta(15), mu, s2, dmu, ds2)
    assert np.allclose(u, ru) and np.allclose(du, rdu)
    th = policies.Thompson(gp, None, None, n=50, rng=3)
    assert th(Xc).shape == (20,)
    assert np.array_equal(recommenders.best_incumbent(gp, None, list(X)), X[np.argmax(gp.predict(X)[0])])


# ---- BASELINE config 1 on CPU through the host glue ---------------------------
def _branin(x):
    x = np.array(x, ndmin=2)
    y = (x[:, 1] - (5.1 / (4 * np.pi ** 2)) * x[:, 0] ** 2 + 5 * x[:, 0] / np.pi - 6) ** 2
    y += 10 * (1 - 1 / (8 * np.pi)) * np.cos(x[:, 0]) + 10
    return float(-np.squeeze(y / 10.0))


def test_bayesopt_branin_trace_matches_golden(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "bayesopt_branin_ei_20.npz"))
    bounds = g["bounds"]
    model = GPOracle(1e-6, 10.0, 0.25 * (bounds[:, 1] - bounds[:, 0]), -5.0, "se")
    log = str(tmp_path / "bo.pkl")
    xbest, out_model, info = pybo_b200.solve_bayesopt(_branin, bounds, model=model, niter=19, policy="ei",
                                                      solver="lbfgs", recommender="latent", rng=0, log=log)
    assert info.x.shape == (20, 2) and info.xbest.shape == (19, 2)
    assert np.allclose(info.x[0], [2.5, 7.5])            # the mid-point query (bayesopt.py:253-258)
    assert model.ndata == 0 and out_model.ndata == 20     # caller's model untouched (bayesopt.py:249)
    # first iterations reproduce the stored trace tightly; BO is chaotic afterwards
    assert np.allclose(info.y[:6], g["y"][:6], rtol=1e-6, atol=1e-8)
    assert info.y.max() > -0.1                            # Branin/10 optimum is -0.0398
    # resume from the checkpoint: nothing left to do, same answer, no new evaluations
    calls = []
    xb2, _, info2 = pybo_b200.solve_bayesopt(lambda x: calls.append(1) or 0.0, bounds, niter=19, log=log)
    assert not calls and np.array_equal(info2.y, info.y) and np.array_equal(xb2, xbest)


def test_model_pickles_without_device_state():
    gp = models.make_gp(1e-6, 1.0, [0.25, 0.5], 0.2, kernel="matern52")
    gp.params["kern.rho"].set_prior("lognormal", 0.0, 1.0)
    gp.add_data(np.random.rand(4, 2), np.random.rand(4))
    gp2 = pickle.loads(pickle.dumps(gp))
    assert gp2.ndata == 4 and gp2.kernel == "matern52" and gp2._fit is None
    assert gp2.params["kern.rho"].prior.name == "lognormal" and gp2.params["kern.rho"].value == 1.0
    c = gp.copy()
    c.add_data(np.random.rand(1, 2), [0.3])
    assert gp.ndata == 4 and c.ndata == 5
    mu, s2 = models.make_gp(1e-6, 2.0, [0.3], 0.7).predict(np.random.rand(5, 1))
    assert np.allclose(mu, 0.7) and np.allclose(s2, 2.0)


def test_priors_and_theta_roundtrip():
    gp = models.make_gp(1e-3, 2.0, [0.3, 0.6], -0.4)
    th = gp.get_theta()
    gp.set_theta(th)
    assert np.allclose([gp.sn2, gp.rho, gp.bias], [1e-3, 2.0, -0.4]) and np.allclose(gp.ell, [0.3, 0.6])
    gp.params["kern.ell"].set_prior("uniform", [0.01, 0.01], [1.0, 1.0])
    assert np.isfinite(gp.logprior())
    gp.ell = np.array([0.3, 5.0])
    assert gp.logprior() == -np.inf
    with pytest.raises(ValueError):
        gp.params["kern.rho"].set_prior("cauchy", 1.0)
