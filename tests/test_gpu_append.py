"""GPU tests of the incremental refit (bo_append): appending observations to a fitted
handle must give the factors, posterior and scores of a fresh fit on all the data (oracle =
refactorisation in LAPACK), through the C ABI and through `model.add_data`."""

import numpy as np
import pytest
from scipy.stats import qmc

from conftest import rel_err
from oracle import GPOracle, MixtureOracle

pytestmark = pytest.mark.gpu


def data(n, d, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    return X, y


def oracle(kernel, X, y, ell, rho, sn2, bias):
    gp = GPOracle(sn2, rho, ell, bias, kernel)
    gp.add_data(X, y)
    return gp


@pytest.mark.parametrize("kernel,n0,k,d", [("se", 1, 5, 2), ("se", 40, 7, 3), ("matern52", 250, 6, 5),
                                           ("se", 1000, 24, 8)])
def test_append_matches_fresh_fit(ctx, kernel, n0, k, d):
    X, y = data(n0 + k, d, seed=n0)
    rho, bias, ell = float(y.max() - y.min()) + 0.5, float(y.mean()), 0.3 * np.ones(d)
    ctx.fit(kernel, X[:n0], y[:n0], ell[None], [rho], [1e-6], [bias])
    assert ctx.capacity() == -(-n0 // 128) * 128
    ctx.append(X[n0:n0 + 1], y[n0:n0 + 1])                  # one point
    ctx.append(X[n0 + 1:n0 + 4], y[n0 + 1:n0 + 4])          # a batch of three
    for i in range(n0 + 4, n0 + k):
        ctx.append(X[i], y[i:i + 1])
    assert ctx.n == n0 + k
    gp = oracle(kernel, X, y, ell, rho, 1e-6, bias)
    L, W = ctx.factor("L"), ctx.factor("W")
    assert L.shape == (n0 + k, n0 + k)
    assert np.max(np.abs(L - gp.L)) < 1e-9 * np.max(np.abs(gp.L))
    assert np.max(np.abs(np.triu(W, 1))) == 0.0
    assert np.max(np.abs(W @ gp.L - np.eye(n0 + k))) < 1e-7
    assert rel_err(ctx.factor("alpha"), gp.alpha, 1e-9) < 1e-6
    assert rel_err(ctx.factor("beta"), gp.beta, 1e-9) < 1e-5
    assert abs(ctx.loglik()[0] - gp.loglikelihood()) < 1e-8 * abs(gp.loglikelihood())
    Xc = qmc.Sobol(d=d, scramble=False).random_base2(9)
    mu, s2, dmu, ds2 = ctx.predict(Xc, grad=True)
    rmu, rs2, rdmu, rds2 = gp.predict(Xc, grad=True)
    assert rel_err(mu, rmu) < 1e-6 and rel_err(s2, rs2, 1e-9) < 1e-6
    assert rel_err(dmu, rdmu, 1e-9) < 1e-5 and rel_err(ds2, rds2, 1e-9) < 1e-5
    target = float(gp.predict(X)[0].max())
    val, _, best = ctx.score(1, target, Xc, want_best=True)
    ref = gp.get_improvement(target, Xc)
    assert rel_err(val, ref, 1e-9) < 1e-6 and best[1] == int(np.argmax(ref))
    # a fresh device fit on the same data agrees to rounding
    from pybo_b200 import _lib
    fresh = _lib.Context(0)
    fresh.fit(kernel, X, y, ell[None], [rho], [1e-6], [bias])
    fval, _, fbest = fresh.score(1, target, Xc, want_best=True)
    assert rel_err(val, fval, 1e-9) < 1e-7 and best[1] == fbest[1]
    fresh.close()


def test_append_capacity_and_errors(ctx):
    X, y = data(130, 2, seed=5)
    ell = 0.3 * np.ones(2)
    from pybo_b200 import _lib
    with pytest.raises(_lib.BackendError):
        ctx.append(X[:1], y[:1])                            # before any fit
    ctx.fit("se", X[:126], y[:126], ell[None], [1.0], [1e-6], [0.0])
    ctx.append(X[126:128], y[126:128])
    assert ctx.n == 128 == ctx.capacity()
    with pytest.raises(_lib.BackendError):
        ctx.append(X[128:129], y[128:129])                  # full: the caller refits
    assert ctx.n == 128
    with pytest.raises(ValueError):
        ctx.append(np.zeros((1, 3)), np.zeros(1))
    mu, s2 = ctx.predict(X)                                  # the failed calls left the fit intact
    rmu, rs2 = oracle("se", X[:128], y[:128], ell, 1.0, 1e-6, 0.0).predict(X)
    assert rel_err(mu, rmu) < 1e-6 and rel_err(s2, rs2, 1e-9) < 1e-6


def test_append_mixture_and_int8_slices(ctx):
    """S = 3 hyper-samples, and the int8 slice planes follow the appended rows."""
    n0, k, d = 500, 12, 8
    X, y = data(n0 + k, d, seed=9)
    rng = np.random.RandomState(1)
    ell = 0.25 * np.exp(0.1 * rng.randn(3, d))
    rho = (float(y.max() - y.min())) * np.exp(0.1 * rng.randn(3))
    sn2 = 1e-6 * np.ones(3)
    bias = float(y.mean()) * np.ones(3)
    ctx.fit("se", X[:n0], y[:n0], ell, rho, sn2, bias)
    ctx.set_precision(1, 1e-8)
    Xc = qmc.Sobol(d=d, scramble=False).random_base2(11)
    ctx.score(1, 0.0, Xc)                                   # builds the slice planes of W before the appends
    for i in range(n0, n0 + k):
        ctx.append(X[i], y[i:i + 1])
    mix = MixtureOracle([oracle("se", X, y, ell[s], rho[s], sn2[s], bias[s]) for s in range(3)])
    target = float(mix.predict(X)[0].max())
    ref = mix.get_improvement(target, Xc)
    v8, _, b8 = ctx.score(1, target, Xc, want_best=True)
    assert rel_err(v8, ref, 1e-9) < 1e-6 and b8[1] == int(np.argmax(ref))
    ctx.set_precision(0)
    v64, _, b64 = ctx.score(1, target, Xc, want_best=True)
    assert rel_err(v64, ref, 1e-9) < 1e-6 and b64[1] == b8[1]
    mu, s2 = ctx.predict(Xc)
    rmu, rs2 = mix.predict(Xc)
    assert rel_err(mu, rmu) < 1e-6 and rel_err(s2, rs2, 1e-9) < 1e-6


def test_model_add_data_appends_in_place_only_when_unshared():
    from pybo_b200 import models
    X, y = data(60, 3, seed=2)
    gp = models.make_gp(1e-6, 2.0, 0.3 * np.ones(3), 0.1)
    gp.add_data(X[:50], y[:50])
    gp.predict(X[:5])
    fit_id = id(gp._fit)                                    # (no strong reference: that would count as sharing)
    gp.add_data(X[50], y[50])                               # sole owner: appended on the device
    assert id(gp._fit) == fit_id and gp._fit.ctx.n == 51
    held = gp.copy()                                        # a policy's private copy shares the handle
    gp.add_data(X[51:53], y[51:53])
    assert gp._fit is None and id(held._fit) == fit_id and held.ndata == 51 and held._fit.ctx.n == 51
    ref = oracle("se", X[:53], y[:53], 0.3 * np.ones(3), 2.0, 1e-6, 0.1)
    Xc = qmc.Sobol(d=3, scramble=False).random_base2(8)
    mu, s2 = gp.predict(Xc)
    rmu, rs2 = ref.predict(Xc)
    assert rel_err(mu, rmu) < 1e-6 and rel_err(s2, rs2, 1e-9) < 1e-6
    hmu, _ = held.predict(Xc)                               # the copy still answers for its 51 points
    r51 = oracle("se", X[:51], y[:51], 0.3 * np.ones(3), 2.0, 1e-6, 0.1)
    assert rel_err(hmu, r51.predict(Xc)[0]) < 1e-6
    del held
    gp.add_data(X[53:60], y[53:60])                         # unshared again after the refit
    assert gp._fit is not None and gp._fit.ctx.n == 60
    gp.incremental = False
    gp.add_data(X[0] + 0.01, y[0])
    assert gp._fit is None
