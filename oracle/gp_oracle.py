"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- parity unpinned.

Float64 NumPy/SciPy/LAPACK restatement of the GP Bayesian-optimisation hot path
that mwhoffman/pybo reaches through its `model` duck type.  The model arithmetic
is not in /root/reference (it lives in the absent, un-pinned `reggie`
package, reference `requirements.txt:8`), so each function below cites the
reference *call site* it serves and SURVEY.md section 8a-math, which is the spec.

LAPACK calls are the ones the reference reaches through reggie/SciPy:
`scipy.linalg.cholesky` (dpotrf), `scipy.linalg.solve_triangular` (dtrtrs),
NumPy matmul (dgemm) and `scipy.special.ndtr`.
"""

import numpy as np
import scipy.linalg as sla
import scipy.special as sps

KERNELS = ("se", "matern52")

# Variance floor applied to the latent predictive variance before any sqrt.
# The reference does not pin one (unverifiable reggie detail); the floor only
# matters where rho - |v|^2 cancels to <= 0 in floating point.
S2_FLOOR = 1e-300

_SQRT5 = np.sqrt(5.0)
_INV_SQRT_2PI = 1.0 / np.sqrt(2.0 * np.pi)


# ----------------------------------------------------------------------------
# kernels (scaled squared distance, SE/RBF and Matern-5/2 ARD)
# ----------------------------------------------------------------------------

def _as2d(X):
    return np.array(X, dtype=np.float64, ndmin=2)


def scaled_sqdist(A, B, ell):
    """D_ij = sum_k ((a_ik - b_jk)/ell_k)^2, evaluated as explicit differences
    (no |a|^2+|b|^2-2ab expansion, so no cancellation)."""
    A = _as2d(A) / ell
    B = _as2d(B) / ell
    diff = A[:, None, :] - B[None, :, :]
    return np.einsum("ijk,ijk->ij", diff, diff)


def kernel_matrix(kind, A, B, ell, rho):
    """k(A, B) -> (len(A), len(B)).  rho is the signal *variance* (kern.rho,
    reference bayesopt.py:99,105), ell the ARD length scales (bayesopt.py:101)."""
    D = scaled_sqdist(A, B, np.asarray(ell, dtype=np.float64))
    if kind == "se":
        return rho * np.exp(-0.5 * D)
    if kind == "matern52":
        r = _SQRT5 * np.sqrt(D)
        return rho * (1.0 + r + r * r / 3.0) * np.exp(-r)
    raise ValueError("unknown kernel %r" % (kind,))


def kernel_gradx(kind, Xc, X, ell, rho):
    """d k(xc_m, x_j) / d xc_m  -> (M, n, d)   (SURVEY 8a-math gradients)."""
    ell = np.asarray(ell, dtype=np.float64)
    Xc = _as2d(Xc)
    X = _as2d(X)
    diff = (Xc[:, None, :] - X[None, :, :]) / (ell * ell)     # (M, n, d)
    D = scaled_sqdist(Xc, X, ell)
    if kind == "se":
        return -(rho * np.exp(-0.5 * D))[:, :, None] * diff
    if kind == "matern52":
        r = _SQRT5 * np.sqrt(D)
        return -(rho * (5.0 / 3.0) * (1.0 + r) * np.exp(-r))[:, :, None] * diff
    raise ValueError("unknown kernel %r" % (kind,))


# ----------------------------------------------------------------------------
# acquisition closed forms on (mu, s2)
# ----------------------------------------------------------------------------

def _pdf(z):
    return _INV_SQRT_2PI * np.exp(-0.5 * z * z)


def ei_from_moments(target, mu, s2, dmu=None, ds2=None):
    """Expected improvement over `target` for maximisation on the latent f
    (model.get_improvement, reference policies/simple.py:25).
    EI = (mu-t) Phi(z) + s phi(z);  dEI = Phi(z) dmu + phi(z)/(2 s) ds2."""
    s2 = np.maximum(s2, S2_FLOOR)
    s = np.sqrt(s2)
    d = mu - target
    z = d / s
    cdf = sps.ndtr(z)
    pdf = _pdf(z)
    ei = d * cdf + s * pdf
    if dmu is None:
        return ei
    return ei, cdf[:, None] * dmu + (0.5 * pdf / s)[:, None] * ds2


def pi_from_moments(target, mu, s2, dmu=None, ds2=None):
    """Probability of improvement P(f > target) = Phi(z)
    (model.get_tail, reference policies/simple.py:39).
    dPI = phi(z)/s (dmu - z/(2 s) ds2)."""
    s2 = np.maximum(s2, S2_FLOOR)
    s = np.sqrt(s2)
    z = (mu - target) / s
    cdf = sps.ndtr(z)
    if dmu is None:
        return cdf
    pdf = _pdf(z)
    return cdf, (pdf / s)[:, None] * (dmu - (0.5 * z / s)[:, None] * ds2)


def ucb_beta(nobs, delta=0.1, xi=0.2):
    """beta of reference policies/simple.py:58-66.  NOTE d = len(X) there is the
    number of *observations*, not the input dimension."""
    a = xi * 2 * np.log(np.pi ** 2 / 3 / delta)
    b = xi * (4 + nobs)
    return a + b * np.log(nobs + 1)


def ucb_index(beta, mu, s2, dmu=None, ds2=None):
    """Reference policies/simple.py:67-72 on given moments."""
    if dmu is None:
        return mu + np.sqrt(beta * s2)
    return (mu + np.sqrt(beta * s2),
            dmu + 0.5 * np.sqrt(beta / s2[:, None]) * ds2)


# ----------------------------------------------------------------------------
# exact GP (the reggie `make_gp(sn2, rho, ell, bias)` model, bayesopt.py:105)
# ----------------------------------------------------------------------------

class GPOracle(object):
    """Exact GP regression with a constant mean and a stationary ARD kernel.

    Duck type observed at the reference call sites (SURVEY 8a row M*):
    add_data (bayesopt.py:114,258,269), copy (simple.py:20,34,57),
    predict (simple.py:21,64; recommenders.py:22,24,34),
    get_improvement (simple.py:25), get_tail (simple.py:39),
    sample_f (simple.py:48).
    """

    def __init__(self, sn2, rho, ell, bias=0.0, kernel="se"):
        if kernel not in KERNELS:
            raise ValueError("unknown kernel %r" % (kernel,))
        self.kernel = kernel
        self.sn2 = float(sn2)
        self.rho = float(rho)
        self.ell = np.array(ell, dtype=np.float64, ndmin=1)
        self.bias = float(bias)
        self.X = np.zeros((0, len(self.ell)))
        self.Y = np.zeros((0,))
        self.L = None
        self.alpha = None      # L^-1 (y - bias)
        self.beta = None       # K^-1 (y - bias) = L^-T alpha

    # -- data ---------------------------------------------------------------
    @property
    def ndata(self):
        return len(self.Y)

    def copy(self):
        new = GPOracle(self.sn2, self.rho, self.ell.copy(), self.bias, self.kernel)
        new.X, new.Y = self.X.copy(), self.Y.copy()
        new.L, new.alpha, new.beta = self.L, self.alpha, self.beta
        return new

    def add_data(self, X, Y):
        X = np.array(X, dtype=np.float64, ndmin=2)
        Y = np.array(Y, dtype=np.float64, ndmin=1)
        if X.shape[0] != Y.shape[0]:
            raise ValueError("X and Y must have the same number of rows")
        self.X = np.concatenate([self.X.reshape(-1, X.shape[1]), X], axis=0)
        self.Y = np.concatenate([self.Y, Y])
        self._refit()

    def _refit(self):
        """K = k(X,X) + sn2 I;  L = chol(K);  alpha = L^-1 (y - bias)."""
        K = kernel_matrix(self.kernel, self.X, self.X, self.ell, self.rho)
        K[np.diag_indices_from(K)] += self.sn2
        self.L = sla.cholesky(K, lower=True)
        r = self.Y - self.bias
        self.alpha = sla.solve_triangular(self.L, r, lower=True)
        self.beta = sla.solve_triangular(self.L, self.alpha, lower=True, trans=1)

    def gram(self):
        K = kernel_matrix(self.kernel, self.X, self.X, self.ell, self.rho)
        K[np.diag_indices_from(K)] += self.sn2
        return K

    def loglikelihood(self):
        """log marginal likelihood (what each MCMC step evaluates)."""
        n = self.ndata
        return (-0.5 * float(self.alpha @ self.alpha)
                - float(np.sum(np.log(np.diag(self.L))))
                - 0.5 * n * np.log(2.0 * np.pi))

    # -- posterior ----------------------------------------------------------
    def predict(self, X, grad=False):
        """Latent posterior (no noise added): mu = bias + V^T alpha,
        s2 = rho - colsum(V^2), V = L^-1 k(Xobs, X)."""
        X = _as2d(X)
        M = X.shape[0]
        if self.ndata == 0:
            mu = np.full(M, self.bias)
            s2 = np.full(M, self.rho)
            if not grad:
                return mu, s2
            return mu, s2, np.zeros_like(X), np.zeros_like(X)
        Ks = kernel_matrix(self.kernel, self.X, X, self.ell, self.rho)   # (n, M)
        V = sla.solve_triangular(self.L, Ks, lower=True)
        mu = self.bias + V.T @ self.alpha
        s2 = self.rho - np.sum(V * V, axis=0)
        if not grad:
            return mu, s2
        dK = kernel_gradx(self.kernel, X, self.X, self.ell, self.rho)    # (M, n, d)
        U = sla.solve_triangular(self.L, V, lower=True, trans=1)         # (n, M)
        dmu = np.einsum("mjk,j->mk", dK, self.beta)
        ds2 = -2.0 * np.einsum("mjk,jm->mk", dK, U)
        return mu, s2, dmu, ds2

    def get_improvement(self, target, X, grad=False):
        post = self.predict(X, grad=grad)
        return ei_from_moments(target, *post)

    def get_tail(self, target, X, grad=False):
        post = self.predict(X, grad=grad)
        return pi_from_moments(target, *post)

    def sample_f(self, n, rng=None):
        return FourierSampleOracle(self, n, rng)


# ----------------------------------------------------------------------------
# hyper-sample mixture (the reggie `MCMC` meta-model, bayesopt.py:115)
# ----------------------------------------------------------------------------

class MixtureOracle(object):
    """Equal-weight mixture over S GPs that share the data but not the hypers
    (SURVEY 8a-math 'Hyper-sample mixture')."""

    def __init__(self, models):
        self.models = list(models)

    def copy(self):
        return MixtureOracle([m.copy() for m in self.models])

    def add_data(self, X, Y):
        for m in self.models:
            m.add_data(X, Y)

    @property
    def ndata(self):
        return self.models[0].ndata

    def predict(self, X, grad=False):
        parts = [m.predict(X, grad=grad) for m in self.models]
        mus = np.array([p[0] for p in parts])
        s2s = np.array([p[1] for p in parts])
        mu = mus.mean(axis=0)
        s2 = (s2s + (mus - mu) ** 2).mean(axis=0)
        if not grad:
            return mu, s2
        dmus = np.array([p[2] for p in parts])
        ds2s = np.array([p[3] for p in parts])
        dmu = dmus.mean(axis=0)
        ds2 = (ds2s + 2.0 * (mus - mu)[:, :, None] * (dmus - dmu)).mean(axis=0)
        return mu, s2, dmu, ds2

    def _mean_of(self, name, target, X, grad):
        parts = [getattr(m, name)(target, X, grad) for m in self.models]
        if not grad:
            return np.mean(parts, axis=0)
        return (np.mean([p[0] for p in parts], axis=0),
                np.mean([p[1] for p in parts], axis=0))

    def get_improvement(self, target, X, grad=False):
        return self._mean_of("get_improvement", target, X, grad)

    def get_tail(self, target, X, grad=False):
        return self._mean_of("get_tail", target, X, grad)

    def sample_f(self, n, rng=None):
        rng = rng if isinstance(rng, np.random.RandomState) else np.random.RandomState(rng)
        return self.models[rng.randint(len(self.models))].sample_f(n, rng)


# ----------------------------------------------------------------------------
# Thompson: one posterior function draw in weight space
# ----------------------------------------------------------------------------

def sample_spectrum(kernel, ell, m, rng):
    """Draw m spectral frequencies of the stationary kernel (rows of W).
    SE: N(0, diag(1/ell^2)).  Matern-nu: the same Gaussian divided by
    sqrt(G), G ~ Gamma(nu, 1/nu)  (multivariate Student-t with 2 nu dof)."""
    d = len(ell)
    W = rng.randn(m, d) / ell
    if kernel == "matern52":
        nu = 2.5
        W = W / np.sqrt(rng.gamma(nu, 1.0 / nu, size=(m, 1)))
    return W


class FourierSampleOracle(object):
    """f(x) = bias + phi(x)^T theta with phi(x) = sqrt(2 rho/m) cos(W x + b)
    (SURVEY 8a-math 'Thompson'; serves policies/simple.py:48 `.get`).

    RNG call order (shared with the product host code so both see the same
    draw): spectrum (randn, then gamma for Matern), phases rand(m), then the
    posterior noise randn(m)."""

    def __init__(self, gp, m, rng=None):
        rng = rng if isinstance(rng, np.random.RandomState) else np.random.RandomState(rng)
        self.m = int(m)
        self.bias = gp.bias
        self.scale = np.sqrt(2.0 * gp.rho / self.m)
        self.W = sample_spectrum(gp.kernel, gp.ell, self.m, rng)
        self.b = rng.rand(self.m) * 2.0 * np.pi
        if gp.ndata > 0:
            Phi = self.features(gp.X)
            A = Phi.T @ Phi
            A[np.diag_indices_from(A)] += gp.sn2
            L = sla.cholesky(A, lower=True)
            rhs = Phi.T @ (gp.Y - gp.bias)
            mean = sla.cho_solve((L, True), rhs)
            noise = rng.randn(self.m)
            self.theta = mean + np.sqrt(gp.sn2) * sla.solve_triangular(
                L, noise, lower=True, trans=1)
        else:
            self.theta = rng.randn(self.m)

    def features(self, X):
        X = _as2d(X)
        return self.scale * np.cos(X @ self.W.T + self.b)

    def get(self, X, grad=False):
        X = _as2d(X)
        arg = X @ self.W.T + self.b
        F = self.bias + (self.scale * np.cos(arg)) @ self.theta
        if not grad:
            return F
        G = -(self.scale * np.sin(arg) * self.theta) @ self.W
        return F, G

    __call__ = get


def thompson_batch_oracle(gp, m, ndraw, rng=None, shared_basis=True):
    """ndraw weight-space posterior draws (the batched form of `sample_f`, BASELINE config 4).
    Returns (W, b, theta, scale): W (nW, m, d), b (nW, m), theta (ndraw, m); nW = 1 with a shared basis
    (RNG order: spectrum, phases, then randn(ndraw, m)), else nW = ndraw with the RNG order of ndraw
    successive FourierSampleOracle constructions."""
    rng = rng if isinstance(rng, np.random.RandomState) else np.random.RandomState(rng)
    scale = np.sqrt(2.0 * gp.rho / m)
    resid = gp.Y - gp.bias

    def solve(W, b, noise):                       # noise: (R, m)
        Phi = scale * np.cos(gp.X @ W.T + b)
        A = Phi.T @ Phi
        A[np.diag_indices_from(A)] += gp.sn2
        L = sla.cholesky(A, lower=True)
        mean = sla.cho_solve((L, True), Phi.T @ resid)
        return mean[None, :] + np.sqrt(gp.sn2) * sla.solve_triangular(L, noise.T, lower=True, trans=1).T

    if shared_basis:
        W = sample_spectrum(gp.kernel, gp.ell, m, rng)
        b = rng.rand(m) * 2.0 * np.pi
        theta = solve(W, b, rng.randn(ndraw, m))
        return W[None], b[None], theta, scale
    Ws, bs, th = [], [], []
    for _ in range(ndraw):
        W = sample_spectrum(gp.kernel, gp.ell, m, rng)
        b = rng.rand(m) * 2.0 * np.pi
        th.append(solve(W, b, rng.randn(1, m))[0])
        Ws.append(W)
        bs.append(b)
    return np.array(Ws), np.array(bs), np.array(th), scale


# ----------------------------------------------------------------------------
# CPU-baseline variant of the scoring pass (bench.py only): identical algebra, but the scaled squared
# distance goes through one dgemm (|a|^2 + |b|^2 - 2 a.b) instead of the (M, n, d) broadcast, and the
# triangular solve / reductions stay in LAPACK / BLAS -- the fastest honest NumPy/SciPy form of the
# reference path, so that the GPU/CPU ratio is not inflated by a slow distance kernel.  The checker
# (`GPOracle.predict`) keeps the cancellation-free differences.
# ----------------------------------------------------------------------------

def predict_fast(gp, X):
    X = _as2d(X)
    A = gp.X / gp.ell
    B = X / gp.ell
    D = (A * A).sum(axis=1)[:, None] + (B * B).sum(axis=1)[None, :] - 2.0 * (A @ B.T)
    np.maximum(D, 0.0, out=D)
    if gp.kernel == "se":
        np.multiply(D, -0.5, out=D)
        np.exp(D, out=D)
        D *= gp.rho
    else:
        r = np.sqrt(5.0 * D)
        D = gp.rho * (1.0 + r + r * r / 3.0) * np.exp(-r)
    V = sla.solve_triangular(gp.L, D, lower=True, overwrite_b=True, check_finite=False)
    mu = gp.bias + gp.alpha @ V
    s2 = gp.rho - np.einsum("ij,ij->j", V, V)
    return mu, s2
