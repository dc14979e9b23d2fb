"""Generate the committed golden vectors in tests/golden/.

The reference cannot be imported in this environment (Python-2-only, and its
GP arithmetic lives in the absent `reggie` package -- SURVEY.md F2-F4), so the
vectors are produced by the float64 oracle (`oracle/gp_oracle.py`) and, where an
independent implementation exists, cross-checked against it before writing:
scikit-learn's GaussianProcessRegressor for mu / sigma, scipy.stats.norm for
EI / PI.  Run from the repo root:  python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np
from scipy.stats import norm, qmc
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import GPOracle, MixtureOracle, ucb_beta, ucb_index  # noqa: E402
import pybo_b200  # noqa: E402  (host glue only; no GPU needed here)
from pybo_b200 import policies, solvers  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def synth(n, d, seed, kernel):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    ell = 0.25 * np.ones(d)
    rho = float(y.max() - y.min())
    gp = GPOracle(1e-6, rho, ell, float(y.mean()), kernel)
    gp.add_data(X, y)
    return gp


def sklearn_check(gp, Xc, mu, s2):
    base = RBF(gp.ell) if gp.kernel == "se" else Matern(gp.ell, nu=2.5)
    skl = GaussianProcessRegressor(kernel=ConstantKernel(gp.rho) * base, alpha=gp.sn2, optimizer=None)
    skl.fit(gp.X, gp.Y - gp.bias)
    m, s = skl.predict(Xc, return_std=True)
    assert np.max(np.abs(m + gp.bias - mu)) < 1e-7 * max(1.0, np.max(np.abs(mu)))
    assert np.max(np.abs(s ** 2 - s2)) < 1e-6 * gp.rho


def posterior_case(name, n, d, M, kernel, seed):
    gp = synth(n, d, seed, kernel)
    Xc = qmc.Sobol(d=d, scramble=False).random(M)
    mu, s2, dmu, ds2 = gp.predict(Xc, grad=True)
    sklearn_check(gp, Xc, mu, s2)
    target = float(gp.predict(gp.X)[0].max())
    ei, dei = gp.get_improvement(target, Xc, grad=True)
    pi, dpi = gp.get_tail(target + 0.05, Xc, grad=True)
    s = np.sqrt(np.maximum(s2, 1e-300))
    z = (mu - target) / s
    assert np.allclose(ei, (mu - target) * norm.cdf(z) + s * norm.pdf(z), rtol=1e-10, atol=1e-300)
    beta = ucb_beta(n)
    ucb, ducb = ucb_index(beta, mu, s2, dmu, ds2)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kernel=kernel, X=gp.X, Y=gp.Y, ell=gp.ell,
                        rho=gp.rho, sn2=gp.sn2, bias=gp.bias, Xc=Xc, mu=mu, s2=s2, dmu=dmu, ds2=ds2,
                        target=target, ei=ei, dei=dei, pi=pi, dpi=dpi, beta=beta, ucb=ucb, ducb=ducb,
                        L=gp.L if n <= 64 else gp.L[:8, :8], alpha=gp.alpha, loglik=gp.loglikelihood())


def mixture_case(name, n, d, M, S, seed):
    rng = np.random.RandomState(seed)
    base = synth(n, d, seed, "se")
    gps = []
    for _ in range(S):
        g = GPOracle(base.sn2 * np.exp(0.3 * rng.randn()), base.rho * np.exp(0.2 * rng.randn()),
                     base.ell * np.exp(0.2 * rng.randn(d)), base.bias + 0.05 * rng.randn(), "se")
        g.add_data(base.X, base.Y)
        gps.append(g)
    mix = MixtureOracle(gps)
    Xc = qmc.Sobol(d=d, scramble=False).random(M)
    mu, s2, dmu, ds2 = mix.predict(Xc, grad=True)
    target = float(mix.predict(base.X)[0].max())
    ei, dei = mix.get_improvement(target, Xc, grad=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=base.X, Y=base.Y,
                        ell=np.array([g.ell for g in gps]), rho=np.array([g.rho for g in gps]),
                        sn2=np.array([g.sn2 for g in gps]), bias=np.array([g.bias for g in gps]),
                        Xc=Xc, mu=mu, s2=s2, dmu=dmu, ds2=ds2, target=target, ei=ei, dei=dei)


def branin(x):
    """Branin / 10, negated for maximisation (reference demos/animated2.py:23-34)."""
    x = np.array(x, ndmin=2)
    y = (x[:, 1] - (5.1 / (4 * np.pi ** 2)) * x[:, 0] ** 2 + 5 * x[:, 0] / np.pi - 6) ** 2
    y += 10 * (1 - 1 / (8 * np.pi)) * np.cos(x[:, 0]) + 10
    return float(-np.squeeze(y / 10.0))


def bayesopt_trace(name):
    """BASELINE config 1: Branin 2D, EI, 20 observations, fixed seed, CPU oracle model
    driven through the product's host glue (solve_bayesopt / EI / solve_lbfgs / best_latent)."""
    bounds = np.array([[-5, 10.0], [0, 15]])
    model = GPOracle(1e-6, 10.0, 0.25 * (bounds[:, 1] - bounds[:, 0]), -5.0, "se")
    xbest, model, info = pybo_b200.solve_bayesopt(branin, bounds, model=model, niter=19, policy="ei",
                                                  solver="lbfgs", recommender="latent", rng=0)
    assert len(info.y) == 20
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=info.x, y=info.y, xbest=info.xbest, bounds=bounds)


if __name__ == "__main__":
    posterior_case("posterior_se_n40_d2", 40, 2, 64, "se", 0)
    posterior_case("posterior_se_n200_d4", 200, 4, 256, "se", 1)
    posterior_case("posterior_matern_n300_d8", 300, 8, 256, "matern52", 2)
    posterior_case("posterior_se_n130_d3", 130, 3, 100, "se", 3)
    mixture_case("mixture_se_n150_d3_s5", 150, 3, 128, 5, 4)
    bayesopt_trace("bayesopt_branin_ei_20")
    print("golden vectors written to", OUT)
