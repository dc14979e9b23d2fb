"""GPU-backed GP models with the duck type pybo's plugins call.

The reference keeps its model layer in the external `reggie` package
(reference bayesopt.py:18,105,115); this module supplies the same surface --
``make_gp``, ``GP``, ``MCMC`` with ``params[...].set_prior``, ``add_data``,
``copy``, ``predict``, ``get_improvement``, ``get_tail``, ``sample_f`` -- on top
of libbo_b200.so.  All numerics for data-bearing models run on the device;
without the CUDA library these classes raise `BackendError`.
"""

import weakref

import numpy as np

from . import _lib
from .utils import rstate

__all__ = ["make_gp", "GP", "MCMC", "FourierSample", "ThompsonBatch"]

_LOG2PI = np.log(2.0 * np.pi)


# ----------------------------------------------------------------------------
# hyper-parameters and priors (reference bayesopt.py:108-111)
# ----------------------------------------------------------------------------

class _Prior(object):
    def __init__(self, name, *args):
        self.name = name
        self.args = [np.array(a, dtype=float) for a in args]
        if name == "uniform":
            if len(args) != 2:
                raise ValueError("uniform prior takes (a, b)")
        elif name in ("lognormal", "normal"):
            if len(args) != 2:
                raise ValueError("%s prior takes (mu, s2)" % name)
        elif name == "horseshoe":
            if len(args) != 1:
                raise ValueError("horseshoe prior takes (scale,)")
        else:
            raise ValueError("unknown prior %r" % (name,))

    def logp(self, theta):
        theta = np.asarray(theta, dtype=float)
        if self.name == "uniform":
            a, b = self.args
            return 0.0 if (np.all(theta >= a) and np.all(theta <= b)) else -np.inf
        if self.name == "lognormal":
            mu, s2 = self.args
            if np.any(theta <= 0):
                return -np.inf
            lt = np.log(theta)
            return float(np.sum(-0.5 * (lt - mu) ** 2 / s2 - lt - 0.5 * np.log(2 * np.pi * s2)))
        if self.name == "normal":
            mu, s2 = self.args
            return float(np.sum(-0.5 * (theta - mu) ** 2 / s2 - 0.5 * np.log(2 * np.pi * s2)))
        scale, = self.args            # horseshoe (Carvalho et al. bound form)
        if np.any(theta <= 0):
            return -np.inf
        return float(np.sum(np.log(np.log1p(3.0 * (scale / theta) ** 2))))


class _Param(object):
    """One named hyper-parameter block with an optional prior."""

    def __init__(self, owner, attr, positive):
        # weak back-reference: a model and its parameter blocks must not form a cycle, so that dropping the last
        # reference to a `model.copy()` (a policy's private copy) releases its share of the fitted handle at once
        self._owner, self._attr, self.positive = weakref.ref(owner), attr, positive
        self.prior = None

    @property
    def value(self):
        return getattr(self._owner(), self._attr)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_owner"] = None               # re-linked by the owner's __setstate__
        return state

    def set_prior(self, name, *args):
        self.prior = None if name is None else _Prior(name, *args)


# ----------------------------------------------------------------------------
# device-resident factorisation, shared between copies of a model
# ----------------------------------------------------------------------------

class _Fit(object):
    """Immutable fitted state: one libbo_b200 handle holding L, W = L^-1, alpha
    for S hyper-samples.  `model.copy()` (policies/simple.py:20,34,57) shares it,
    `add_data` drops the reference and the next use refits."""

    def __init__(self, kernel, X, Y, ell, rho, sn2, bias, device=None):
        self.ctx = _lib.Context(device)
        self.ctx.fit(kernel, X, Y, ell, rho, sn2, bias)
        self.owners = 0                      # models currently holding this handle (`_Base._hold` / `_drop`)
        self.applied = ("fp64", None)        # precision path last written into the handle


class _Base(object):
    """Shared data handling; subclasses define the hyper-sample arrays."""

    kernel = "se"
    device = None
    _precision = ("fp64", 1e-9)
    incremental = True           # add_data appends to the device factors when it can

    def set_precision(self, path="fp64", tol=1e-7):
        """Arithmetic of the scoring contraction: 'fp64' (FP64 tensor cores, default) or 'int8'
        (error-bounded int8 slices on tcgen05; `tol` is the target absolute error of V = L^-1 k
        relative to sqrt(rho), or an explicit slice count when >= 2).  Gradients, small batches and
        `predict` at fewer than 65 points always use FP64."""
        if path not in ("fp64", "int8"):
            raise ValueError("precision path must be 'fp64' or 'int8'")
        # per-model state, written into the (possibly shared) handle right before each device call: setting the
        # precision of one copy never changes the arithmetic of another model sharing the handle
        self._precision = (path, float(tol))
        return self

    _fit = None

    def _init_data(self, d):
        self._X = np.zeros((0, d))
        self._Y = np.zeros((0,))
        self._drop()

    # -- ownership of the shared fitted handle ---------------------------------------------------------
    def _hold(self, fit):
        """Take a share of `fit` (or of nothing)."""
        self._drop()
        if fit is not None:
            fit.owners += 1
        self._fit = fit

    def _drop(self):
        fit = self.__dict__.get("_fit")
        if fit is not None:
            fit.owners -= 1
        self._fit = None
        return fit

    def __del__(self):
        try:
            self._drop()
        except Exception:
            pass

    # hyper-sample view: (ell[S,d], rho[S], sn2[S], bias[S])
    def _hypers(self):
        raise NotImplementedError

    @property
    def ndata(self):
        return len(self._Y)

    @property
    def data(self):
        return self._X, self._Y

    def add_data(self, X, Y):
        """Append observations (accepts lists, or one (d,) point and a scalar:
        reference bayesopt.py:114,258,269)."""
        d = self._X.shape[1]
        X = np.array(X, dtype=np.float64, ndmin=2)
        Y = np.array(Y, dtype=np.float64, ndmin=1)
        if X.shape[1] != d or X.shape[0] != Y.shape[0]:
            raise ValueError("add_data: expected (k, %d) inputs and (k,) outputs" % d)
        self._X = np.concatenate([self._X, X], axis=0)
        self._Y = np.concatenate([self._Y, Y])
        self._extend_fit(X, Y)

    def _extend_fit(self, X, Y):
        """Incremental refit: when this model is the only owner of its fitted handle
        (no `copy()` still shares it) and the handle has room, the new rows of L, W, alpha, beta are
        appended on the device in O(n^2) per point (bo_append) instead of refactorising; otherwise
        the handle is dropped and the next use refits."""
        fit = self._drop()
        if fit is None or not self.incremental or len(Y) == 0 or len(Y) > 8:
            return
        if fit.owners > 0:                      # a `copy()` still shares the handle: never mutate it
            return
        try:
            if fit.ctx.n + len(Y) > fit.ctx.capacity():
                return
            fit.ctx.append(X, Y)
        except (np.linalg.LinAlgError, _lib.BackendError):
            return
        self._hold(fit)

    def _ensure_fit(self):
        if self._fit is None:
            ell, rho, sn2, bias = self._hypers()
            self._hold(_Fit(self.kernel, self._X, self._Y, ell, rho, sn2, bias, self.device))
        fit = self._fit
        if fit.applied != self._precision:
            path, tol = self._precision
            fit.ctx.set_precision(1 if path == "int8" else 0, tol)
            fit.applied = self._precision
        return fit.ctx

    # pickling / checkpointing (reference bayesopt.py:39-55): device state is rebuilt lazily
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_fit"] = None
        if "_sampler_ctx" in state:
            state["_sampler_ctx"] = None
        return state

    # -- posterior ------------------------------------------------------------
    def _prior_moments(self, X, grad):
        ell, rho, sn2, bias = self._hypers()
        M = X.shape[0]
        mu = np.full(M, np.mean(bias))
        s2 = np.full(M, np.mean(rho + (bias - np.mean(bias)) ** 2))
        if grad:
            return mu, s2, np.zeros_like(X), np.zeros_like(X)
        return mu, s2

    def predict(self, X, grad=False):
        X = np.array(X, dtype=np.float64, ndmin=2)
        if self.ndata == 0:
            return self._prior_moments(X, grad)
        return self._ensure_fit().predict(X, grad=grad)

    def _acq(self, acq, param, X, grad):
        X = np.array(X, dtype=np.float64, ndmin=2)
        if self.ndata == 0:
            raise ValueError("acquisition functions need at least one observation")
        val, g, _ = self._ensure_fit().score(acq, param, X, grad=grad)
        return (val, g) if grad else val

    def get_improvement(self, target, X, grad=False):
        return self._acq(_lib.ACQ_EI, target, X, grad)

    def get_tail(self, target, X, grad=False):
        return self._acq(_lib.ACQ_PI, target, X, grad)

    def loglikelihood(self):
        if self.ndata == 0:
            return np.zeros(len(self._hypers()[1]))
        return self._ensure_fit().loglik()


class GP(_Base):
    """Exact GP with a constant mean: `make_gp(sn2, rho, ell, bias)`
    (reference bayesopt.py:105, demos/animated2.py:53)."""

    def __init__(self, sn2, rho, ell, bias=0.0, kernel="se", device=None):
        if kernel not in _lib.KERNEL_IDS:
            raise ValueError("unknown kernel %r" % (kernel,))
        self.kernel = kernel
        self.device = device
        self.sn2 = float(sn2)
        self.rho = float(rho)
        self.ell = np.array(ell, dtype=np.float64, ndmin=1).copy()
        self.bias = float(bias)
        self._init_data(len(self.ell))
        self.params = {
            "like.sn2": _Param(self, "sn2", True),
            "kern.rho": _Param(self, "rho", True),
            "kern.ell": _Param(self, "ell", True),
            "mean.bias": _Param(self, "bias", False),
        }

    def _hypers(self):
        return (self.ell[None, :], np.array([self.rho]), np.array([self.sn2]), np.array([self.bias]))

    def copy(self):
        new = GP(self.sn2, self.rho, self.ell, self.bias, self.kernel, self.device)
        new._X, new._Y = self._X, self._Y
        new._hold(self._fit)
        new._precision = self._precision
        for k, p in self.params.items():
            new.params[k].prior = p.prior
        return new

    def __setstate__(self, state):
        self.__dict__.update(state)
        for p in self.params.values():
            p._owner = weakref.ref(self)

    # hyper-parameter vector in the unconstrained space used by the sampler
    def get_theta(self):
        return np.concatenate([[np.log(self.sn2), np.log(self.rho)], np.log(self.ell), [self.bias]])

    def set_theta(self, theta):
        d = len(self.ell)
        self.sn2, self.rho = float(np.exp(theta[0])), float(np.exp(theta[1]))
        self.ell = np.exp(np.asarray(theta[2:2 + d], dtype=float))
        self.bias = float(theta[2 + d])
        self._drop()

    def logprior(self):
        """log prior + log |Jacobian| of the log transform of the positive blocks."""
        lp = 0.0
        for p in self.params.values():
            if p.prior is not None:
                lp += p.prior.logp(p.value)
            if p.positive:
                lp += float(np.sum(np.log(p.value)))
        return lp

    def sample_f(self, n, rng=None):
        """One posterior function draw with n random Fourier features
        (reference policies/simple.py:48)."""
        return FourierSample(self, n, rng)


def make_gp(sn2, rho, ell, bias=0.0, kernel="se", device=None):
    return GP(sn2, rho, ell, bias, kernel=kernel, device=device)


# ----------------------------------------------------------------------------
# hyper-sample mixture (`reggie.MCMC(model, n=10, burn=100, rng)`, bayesopt.py:115)
# ----------------------------------------------------------------------------

def _slice_sample(logp, x0, lp0, rng, width=1.0, max_steps=8):
    """One slice-sampling update along a random direction (step-out + shrink)."""
    direction = rng.randn(len(x0))
    direction *= width / np.sqrt(np.sum(direction ** 2))
    level = lp0 + np.log(rng.rand())
    upper = rng.rand()
    lower = upper - 1.0
    for _ in range(max_steps):
        if logp(x0 + lower * direction) <= level:
            break
        lower -= 1.0
    for _ in range(max_steps):
        if logp(x0 + upper * direction) <= level:
            break
        upper += 1.0
    while True:
        t = lower + (upper - lower) * rng.rand()
        x1 = x0 + t * direction
        lp1 = logp(x1)
        if lp1 > level:
            return x1, lp1
        if t < 0:
            lower = t
        else:
            upper = t
        if upper - lower < 1e-12:
            return x0, lp0


class MCMC(_Base):
    """Equal-weight mixture over n hyper-samples drawn by slice sampling from the
    hyper-posterior of `model`; every device call is batched over the samples."""

    def __init__(self, model, n=10, burn=100, rng=None):
        self._rng = rstate(rng)
        self._proto = model.copy()
        self.kernel = model.kernel
        self.device = model.device
        self._n = int(n)
        self._X, self._Y = model._X.copy(), model._Y.copy()
        self._sampler_ctx = None
        self._thetas = None
        if burn > 0:
            self._resample(burn)
        self._resample(self._n)

    # -- construction from explicit hyper-samples (tests, benchmarks) ----------
    @classmethod
    def from_samples(cls, kernel, ell, rho, sn2, bias, device=None):
        ell = np.array(ell, dtype=np.float64, ndmin=2)
        self = cls.__new__(cls)
        self._rng = rstate(None)
        self._proto = GP(float(np.ravel(sn2)[0]), float(np.ravel(rho)[0]), ell[0], float(np.ravel(bias)[0]),
                         kernel=kernel, device=device)
        self.kernel, self.device = kernel, device
        self._n = ell.shape[0]
        self._init_data(ell.shape[1])
        self._sampler_ctx = None
        self._thetas = np.column_stack([np.log(np.ravel(sn2)), np.log(np.ravel(rho)), np.log(ell), np.ravel(bias)])
        return self

    def _logpost(self, theta):
        """log hyper-posterior at `theta`: host-side priors + the device log marginal
        likelihood (bo_loglik_fit: Gram + Cholesky of the matrix bordered by the residual row -- no W = L^-1, no
        transpose, no scoring state -- on a handle kept for the sampler)."""
        gp = self._proto
        try:
            with np.errstate(over="raise", invalid="raise"):
                gp.set_theta(theta)
                lp = gp.logprior()
            if not np.isfinite(lp):
                return -np.inf
            if len(self._Y) == 0:
                return lp
            if getattr(self, "_sampler_ctx", None) is None:
                self._sampler_ctx = _lib.Context(self.device)
            ctx = self._sampler_ctx
            ll = float(ctx.loglik_fit(self.kernel, self._X, self._Y, gp.ell[None, :], [gp.rho], [gp.sn2], [gp.bias])[0])
        except (np.linalg.LinAlgError, FloatingPointError, OverflowError, ValueError):
            return -np.inf
        val = lp + ll
        return val if np.isfinite(val) else -np.inf

    def _resample(self, count):
        theta = self._thetas[-1].copy() if self._thetas is not None else self._proto.get_theta()
        lp = self._logpost(theta)
        if not np.isfinite(lp):
            raise ValueError("MCMC: initial hyper-parameters have zero posterior density")
        out = []
        for _ in range(count):
            theta, lp = _slice_sample(self._logpost, theta, lp, self._rng)
            out.append(theta.copy())
        self._thetas = np.array(out[-self._n:])
        self._proto._drop()
        self._drop()

    def _hypers(self):
        th = self._thetas
        d = self._X.shape[1]
        return (np.exp(th[:, 2:2 + d]), np.exp(th[:, 1]), np.exp(th[:, 0]), th[:, 2 + d].copy())

    def __len__(self):
        return len(self._thetas)

    def add_data(self, X, Y, resample=True):
        resample = resample and self._proto is not None and any(p.prior is not None for p in self._proto.params.values())
        if resample:
            self._drop()                # new hyper-samples follow: nothing to append to
        _Base.add_data(self, X, Y)
        if resample:
            self._resample(self._n)

    def copy(self):
        new = MCMC.__new__(MCMC)
        state = dict(self.__dict__)
        fit = state.pop("_fit", None)
        new.__dict__.update(state)
        new._hold(fit)
        new._proto = self._proto.copy()
        new._thetas = self._thetas.copy()
        return new

    def sample_f(self, n, rng=None):
        rng = rstate(rng)
        ell, rho, sn2, bias = self._hypers()
        s = rng.randint(len(rho))
        gp = GP(sn2[s], rho[s], ell[s], bias[s], kernel=self.kernel, device=self.device)
        gp._X, gp._Y = self._X, self._Y
        return gp.sample_f(n, rng)


# ----------------------------------------------------------------------------
# Thompson draws
# ----------------------------------------------------------------------------

def _spectrum(kernel, ell, m, rng):
    W = rng.randn(m, len(ell)) / ell
    if kernel == "matern52":
        W = W / np.sqrt(rng.gamma(2.5, 1.0 / 2.5, size=(m, 1)))
    return W


class FourierSample(object):
    """f(x) = bias + sqrt(2 rho / m) cos(W x + b) . theta; `.get(X, grad)` is the
    Thompson index (reference policies/simple.py:48).  The spectral points, phases and the
    standard-normal vector come from `rng` (NumPy stream, in this order); the feature system
    Phi^T Phi + sn2 I, its Cholesky factor and the two solves behind theta run on the device
    (bo_thompson_build), as does every evaluation."""

    def __init__(self, gp, m, rng=None):
        rng = rstate(rng)
        self.m = int(m)
        self.bias = float(gp.bias)
        self.scale = float(np.sqrt(2.0 * gp.rho / self.m))
        self.W = _spectrum(gp.kernel, gp.ell, self.m, rng)
        self.b = rng.rand(self.m) * 2.0 * np.pi
        self.device = gp.device
        self._ctx = None
        noise = rng.randn(self.m)
        if gp.ndata > 0:
            ctx = _lib.Context(self.device)
            self.theta = ctx.thompson_build(gp._X, gp._Y, gp.rho, gp.sn2, gp.bias, self.W[None], self.b[None],
                                            noise[None])[0]
            self._ctx = ctx
        else:
            self.theta = noise

    def _context(self):
        if self._ctx is None:
            ctx = _lib.Context(self.device)
            ctx.thompson_set(self.W[None], self.b[None], self.theta[None], [self.scale], [self.bias])
            self._ctx = ctx
        return self._ctx

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_ctx"] = None
        return state

    def get(self, X, grad=False):
        X = np.array(X, dtype=np.float64, ndmin=2)
        out, g, _ = self._context().thompson_eval(X, grad=grad)
        return (out[0], g[0]) if grad else out[0]

    __call__ = get

    def argmax(self, X):
        X = np.array(X, dtype=np.float64, ndmin=2)
        _, _, (bv, bi) = self._context().thompson_eval(X, want_values=False, want_best=True)
        return float(bv[0]), int(bi[0])


class ThompsonBatch(object):
    """ndraw posterior draws evaluated together (BASELINE config 4: 256 draws x 1M candidates).

    shared_basis=True (default): one random-feature basis (W, b) for all draws, each draw its own theta, so
    evaluation is one dense (M x m) x (m x ndraw) contraction (FP64 tensor cores, or int8 slices on tcgen05).
    shared_basis=False: every draw has its own basis, drawn from `rng` in the order ndraw successive
    `model.sample_f(m, rng)` calls would use -- the literal batched form of policies/simple.py:48.
    Either way the feature systems are built and solved on the device (bo_thompson_build)."""

    def __init__(self, gp, m, ndraw, rng=None, shared_basis=True):
        rng = rstate(rng)
        self.m, self.ndraw = int(m), int(ndraw)
        self.shared_basis = bool(shared_basis)
        self.bias = float(gp.bias)
        self.scale = float(np.sqrt(2.0 * gp.rho / self.m))
        self.device = gp.device
        if self.shared_basis:
            self.W = _spectrum(gp.kernel, gp.ell, self.m, rng)[None]
            self.b = (rng.rand(self.m) * 2.0 * np.pi)[None]
            noise = rng.randn(self.ndraw, self.m)
        else:
            W, b, noise = [], [], []
            for _ in range(self.ndraw):
                W.append(_spectrum(gp.kernel, gp.ell, self.m, rng))
                b.append(rng.rand(self.m) * 2.0 * np.pi)
                noise.append(rng.randn(self.m))
            self.W, self.b, noise = np.array(W), np.array(b), np.array(noise)
        self._ctx = None
        if gp.ndata > 0:
            ctx = _lib.Context(self.device)
            self.theta = ctx.thompson_build(gp._X, gp._Y, gp.rho, gp.sn2, gp.bias, self.W, self.b, noise)
            self._ctx = ctx
            self._applied = ("fp64", None)
        else:
            self.theta = noise

    _precision = ("fp64", 1e-9)
    _applied = ("fp64", None)

    def set_precision(self, path="fp64", tol=1e-8):
        """'fp64': FP64 tensor-core contraction with on-the-fly cosine features (default); 'int8': the
        features and Theta are cut into balanced base-256 int8 slices and contracted on tcgen05 (shared basis,
        batches of >= 1024 candidates; `tol` is the target error relative to a draw's own scale, >= 2 pins the level)."""
        if path not in ("fp64", "int8"):
            raise ValueError("precision path must be 'fp64' or 'int8'")
        self._precision = (path, float(tol))
        return self

    def _context(self):
        if self._ctx is None:
            ctx = _lib.Context(self.device)
            ctx.thompson_set(self.W, self.b, self.theta,
                             np.full(self.ndraw, self.scale), np.full(self.ndraw, self.bias))
            self._ctx = ctx
            self._applied = ("fp64", None)
        if self._applied != self._precision:
            path, tol = self._precision
            self._ctx.set_precision(1 if path == "int8" else 0, tol)
            self._applied = self._precision
        return self._ctx

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_ctx"] = None
        return state

    def get(self, X):
        X = np.array(X, dtype=np.float64, ndmin=2)
        return self._context().thompson_eval(X)[0]

    def argmax(self, X):
        X = np.array(X, dtype=np.float64, ndmin=2)
        _, _, (bv, bi) = self._context().thompson_eval(X, want_values=False, want_best=True)
        return bv, bi
