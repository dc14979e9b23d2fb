"""Hyper-parameter inference (reference bayesopt.py:98-115: priors + MCMC(model, n=10, burn=100)):
the likelihood-only entry point against the oracle, and a statistical test of the slice sampler against
brute-force quadrature of the same log posterior on a two-hyper-parameter toy."""

import numpy as np
import pytest

from oracle import GPOracle

pytestmark = pytest.mark.gpu


def toy(n, d, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(3.0 * X.sum(axis=1)) + 0.1 * rng.randn(n)
    return rng, X, y


@pytest.mark.parametrize("kernel", ["se", "matern52"])
@pytest.mark.parametrize("n,d,S", [(1, 1, 1), (20, 2, 1), (63, 3, 4), (64, 2, 2), (65, 5, 3), (200, 4, 10), (1000, 8, 2)])
def test_loglik_fit_matches_oracle(ctx, kernel, n, d, S):
    rng, X, y = toy(n, d, seed=n + d)
    ell = 0.3 * np.exp(0.2 * rng.randn(S, d))
    rho = (np.ptp(y) + 0.5) * np.exp(0.2 * rng.randn(S))
    sn2 = 1e-2 * np.exp(0.5 * rng.randn(S))
    bias = y.mean() + 0.1 * rng.randn(S)
    got = ctx.loglik_fit(kernel, X, y, ell, rho, sn2, bias)
    want = []
    for s in range(S):
        g = GPOracle(sn2[s], rho[s], ell[s], bias[s], kernel)
        g.add_data(X, y)
        want.append(g.loglikelihood())
    want = np.array(want)
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < 1e-9
    # and it agrees with the scoring-state route (bo_fit + bo_loglik) without disturbing a fitted handle
    ctx.fit(kernel, X, y, ell, rho, sn2, bias)
    full = ctx.loglik()
    again = ctx.loglik_fit(kernel, X, y, ell[::-1].copy(), rho[::-1].copy(), sn2[::-1].copy(), bias[::-1].copy())
    assert np.allclose(full, want, rtol=1e-9, atol=1e-9) and np.allclose(again, want[::-1], rtol=1e-9, atol=1e-9)
    assert np.allclose(ctx.loglik(), full, rtol=0, atol=0)


def test_loglik_fit_not_positive_definite(ctx):
    X = np.zeros((8, 2))                                          # eight copies of one point, no noise
    y = np.arange(8.0)
    with pytest.raises(np.linalg.LinAlgError):
        ctx.loglik_fit("se", X, y, np.ones((1, 2)), [1.0], [0.0], [0.0])


def test_sampler_matches_quadrature_on_two_hypers():
    """The product's slice-sampling update on the product's log posterior (host priors + device bo_loglik_fit),
    restricted to (log rho, log ell) with sn2 and bias held fixed so that the target can be integrated by brute
    force: the chain's mean and variance of both coordinates must agree with quadrature of exp(logpost) on a
    grid.  Then the full four-block sampler with the reference's prior families (bayesopt.py:108-111) must run
    and move."""
    from pybo_b200 import models
    rng, X, y = toy(25, 1, seed=3)
    gp = models.make_gp(0.01, 1.0, [0.3], 0.0)
    gp.params["kern.rho"].set_prior("lognormal", 0.0, 1.0)
    gp.params["kern.ell"].set_prior("lognormal", -1.0, 1.0)
    gp.add_data(X, y)
    mc = models.MCMC.__new__(models.MCMC)
    mc._rng = np.random.RandomState(0)
    mc._proto = gp.copy()
    mc.kernel, mc.device, mc._n = gp.kernel, gp.device, 10
    mc._X, mc._Y = gp._X.copy(), gp._Y.copy()
    mc._sampler_ctx, mc._thetas = None, None
    base = gp.get_theta()                                          # [log sn2, log rho, log ell, bias]

    def logpost2(u):                                               # u = (log rho, log ell); sn2, bias fixed
        th = base.copy()
        th[1], th[2] = u
        return mc._logpost(th)

    # quadrature on a grid wide enough to hold the mass
    g1 = np.linspace(-4.0, 5.0, 91)
    g2 = np.linspace(-4.5, 2.5, 71)
    lp = np.array([[logpost2((a, b)) for b in g2] for a in g1])
    w = np.exp(lp - lp.max())
    w /= w.sum()
    m1, m2 = float((w.sum(1) * g1).sum()), float((w.sum(0) * g2).sum())
    v1, v2 = float((w.sum(1) * (g1 - m1) ** 2).sum()), float((w.sum(0) * (g2 - m2) ** 2).sum())
    # chain in the 2-d subspace with the product's own slice-sampling update
    u = np.array([base[1], base[2]])
    cur = logpost2(u)
    draws = []
    for it in range(3000):
        u, cur = models._slice_sample(logpost2, u, cur, mc._rng)
        if it >= 300:
            draws.append(u.copy())
    draws = np.array(draws)
    # effective sample size is a fraction of 2700; 5-sigma-ish tolerances on the moments
    assert abs(draws[:, 0].mean() - m1) < 0.25 * np.sqrt(v1) + 0.02, (draws[:, 0].mean(), m1, v1)
    assert abs(draws[:, 1].mean() - m2) < 0.25 * np.sqrt(v2) + 0.02, (draws[:, 1].mean(), m2, v2)
    assert 0.6 < draws[:, 0].var() / v1 < 1.6 and 0.6 < draws[:, 1].var() / v2 < 1.6
    # the full sampler (all four blocks, reference priors) runs on the same machinery and mixes
    full = models.make_gp(0.01, 1.0, [0.3], 0.0)
    full.params["like.sn2"].set_prior("horseshoe", 0.1)
    full.params["kern.rho"].set_prior("lognormal", 0.0, 1.0)
    full.params["kern.ell"].set_prior("uniform", 0.01, 5.0)
    full.params["mean.bias"].set_prior("normal", 0.0, 1.0)
    full.add_data(X, y)
    chain = models.MCMC(full, n=10, burn=50, rng=1)
    ell, rho, sn2, bias = chain._hypers()
    assert len(chain) == 10 and np.all(ell > 0.01) and np.all(ell < 5.0) and np.all(sn2 > 0)
    assert len(np.unique(np.round(rho, 12))) > 3                   # the chain moves
