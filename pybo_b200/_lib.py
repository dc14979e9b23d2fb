"""ctypes binding of libbo_b200.so (include/bo_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device
is visible, every entry point that needs the GPU raises `BackendError`.
"""

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "lib", "libbo_b200.so")

BO_OK, BO_ERR_CUDA, BO_ERR_NOT_PD, BO_ERR_ARG, BO_ERR_STATE = 0, 1, 2, 3, 4
KERNEL_IDS = {"se": 0, "matern52": 1}
ACQ_MEAN, ACQ_EI, ACQ_PI, ACQ_UCB = 0, 1, 2, 3
PTR_HOST, PTR_DEVICE, PTR_STAGED = 0, 1, 2

# every symbol include/bo_b200.h declares
EXPORTS = [
    "bo_create", "bo_destroy", "bo_last_error", "bo_stream", "bo_sync", "bo_device_props",
    "bo_fit", "bo_fit_shape", "bo_fit_info", "bo_loglik", "bo_get_factor",
    "bo_score", "bo_predict", "bo_topk", "bo_set_precision", "bo_precision_info",
    "bo_thompson_set", "bo_thompson_eval",
    "bo_cholesky", "bo_gram",
    "bo_profile_enable", "bo_profile_reset", "bo_profile_count", "bo_profile_get",
    "bo_launch_count", "bo_microbench", "bo_ozaki_debug", "bo_append", "bo_fit_capacity", "bo_candidates_sobol",
    "bo_set_rescue", "bo_rescue_info", "bo_ozaki_error_bound", "bo_loglik_fit",
    "bo_thompson_build", "bo_score_incumbent", "bo_incumbent_merge", "bo_thompson_incumbents", "bo_set_option", "bo_tier_info",
]


class BackendError(RuntimeError):
    """The CUDA backend is unavailable or a CUDA call failed."""


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


def _declare(lib):
    vp, i, d, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    sig = {
        "bo_create": (i, [i, C.POINTER(vp)]),
        "bo_destroy": (i, [vp]),
        "bo_last_error": (C.c_char_p, [vp]),
        "bo_stream": (vp, [vp]),
        "bo_sync": (i, [vp]),
        "bo_device_props": (i, [vp, _ip, _ip, _ip, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "bo_fit": (i, [vp, i, i, i, i, vp, vp, vp, vp, vp, vp]),
        "bo_fit_shape": (i, [vp, _ip, _ip, _ip, _ip]),
        "bo_fit_info": (i, [vp, vp]),
        "bo_loglik": (i, [vp, vp]),
        "bo_get_factor": (i, [vp, i, i, vp]),
        "bo_score": (i, [vp, i, d, i64, vp, i, vp, vp, _dp, _lp]),
        "bo_predict": (i, [vp, i64, vp, i, vp, vp, vp, vp]),
        "bo_topk": (i, [vp, i, vp, vp]),
        "bo_set_precision": (i, [vp, i, d]),
        "bo_precision_info": (i, [vp, _ip, _ip]),
        "bo_thompson_set": (i, [vp, i, i, i, i, vp, vp, vp, vp, vp]),
        "bo_thompson_eval": (i, [vp, i64, vp, i, vp, vp, vp, vp]),
        "bo_cholesky": (i, [vp, i, i, vp, i, vp]),
        "bo_gram": (i, [vp, i, i, i, vp, vp, d, d, vp, i]),
        "bo_profile_enable": (i, [vp, i]),
        "bo_profile_reset": (i, [vp]),
        "bo_profile_count": (i, [vp, _ip]),
        "bo_profile_get": (i, [vp, i, C.c_char_p, i, _lp, _dp]),
        "bo_launch_count": (i, [vp, _lp]),
        "bo_microbench": (i, [vp, i, i, _dp]),
        "bo_ozaki_debug": (i, [vp, i, i, i, vp, vp, vp, vp, vp, vp, vp]),
        "bo_append": (i, [vp, i, vp, vp]),
        "bo_fit_capacity": (i, [vp, vp]),
        "bo_candidates_sobol": (i, [vp, i, i, vp, vp, vp, i64, i64, vp, i]),
        "bo_set_rescue": (i, [vp, i, d, d]),
        "bo_rescue_info": (i, [vp, _ip, _lp, _lp]),
        "bo_tier_info": (i, [vp, _ip, _ip, _ip, _lp, _lp]),
        "bo_ozaki_error_bound": (i, [vp, vp]),
        "bo_loglik_fit": (i, [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, vp]),
        "bo_thompson_build": (i, [vp, i, i, vp, vp, d, d, d, i, i, i, vp, vp, vp, vp]),
        "bo_score_incumbent": (i, [vp, i, d, i64, vp, i, vp, i64, C.POINTER(vp)]),
        "bo_set_option": (i, [vp, C.c_char_p, d]),
        "bo_incumbent_merge": (i, [vp, vp, i, i, vp, vp]),
        "bo_thompson_incumbents": (i, [vp, i64, vp, i, i64, C.POINTER(vp), _ip]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


def load():
    """Load the shared library (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise BackendError(
                "libbo_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python -m pybo_b200._build`. There is no CPU fallback." % LIBPATH)
        lib = C.CDLL(LIBPATH)
        _declare(lib)
        _lib = lib
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


_SOBOL_SV = {}


def sobol_directions(d):
    """(sv, bits): direction numbers (d x bits, uint32) of SciPy's unscrambled Sobol generator."""
    if d not in _SOBOL_SV:
        from scipy.stats import qmc
        eng = qmc.Sobol(d=d, scramble=False)
        _SOBOL_SV[d] = (np.ascontiguousarray(eng._sv, dtype=np.uint32), int(eng.bits))
    return _SOBOL_SV[d]


def sobol_points(d, indices, bounds=None):
    """Host restatement of the device generator for a few indices (gray-code XOR of direction numbers)."""
    sv, bits = sobol_directions(d)
    idx = np.asarray(indices, dtype=np.uint64)
    g = idx ^ (idx >> np.uint64(1))
    x = np.zeros((len(idx), d), dtype=np.uint64)
    for b in range(bits):
        on = ((g >> np.uint64(b)) & np.uint64(1)).astype(bool)
        x[on] ^= sv[:, b].astype(np.uint64)
    u = x.astype(np.float64) * 2.0 ** -bits
    if bounds is None:
        return u
    bnd = np.array(bounds, dtype=np.float64, ndmin=2)
    return bnd[:, 0] + (bnd[:, 1] - bnd[:, 0]) * u


def f64(a, ndmin=1):
    """C-contiguous float64 view of `a` with at least `ndmin` dimensions; no copy when `a`
    already is one (candidate grids are large and may live in pinned memory)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    while a.ndim < ndmin:
        a = a[None]
    return a


class Context(object):
    """One `bo_ctx` handle: one device, one stream, one fitted factor set."""

    def __init__(self, device=None):
        lib = load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        rc = lib.bo_create(int(device), C.byref(h))
        if rc != BO_OK or not h.value:
            raise BackendError(
                "bo_create(device=%d) failed (status %d): no usable CUDA device. "
                "pybo_b200 has no CPU fallback." % (device, rc))
        self._h = h
        self._lib = lib
        self.device = int(device)

    # -- plumbing -------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.bo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == BO_OK:
            return
        msg = self._lib.bo_last_error(self._h).decode("utf-8", "replace")
        if rc == BO_ERR_NOT_PD:
            raise np.linalg.LinAlgError(msg)
        if rc == BO_ERR_ARG:
            raise ValueError(msg)
        raise BackendError("libbo_b200 status %d: %s" % (rc, msg))

    def sync(self):
        self._check(self._lib.bo_sync(self._h))

    @property
    def stream(self):
        return self._lib.bo_stream(self._h)

    def device_props(self):
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        l2, hbm = C.c_size_t(), C.c_size_t()
        self._check(self._lib.bo_device_props(self._h, C.byref(sm), C.byref(ma), C.byref(mi),
                                              C.byref(l2), C.byref(hbm)))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), l2_bytes=l2.value, hbm_bytes=hbm.value)

    # -- fit --------------------------------------------------------------------
    def fit(self, kernel, X, y, ell, rho, sn2, bias):
        X = f64(X, 2)
        y = f64(y, 1)
        ell = f64(ell, 2)
        rho, sn2, bias = f64(rho), f64(sn2), f64(bias)
        n, d = X.shape
        S = ell.shape[0]
        if ell.shape[1] != d or len(rho) != S or len(sn2) != S or len(bias) != S or len(y) != n:
            raise ValueError("inconsistent fit shapes")
        self.n, self.d, self.S = n, d, S
        self._check(self._lib.bo_fit(self._h, KERNEL_IDS[kernel], n, d, S, _ptr(X), _ptr(y), _ptr(ell),
                                     _ptr(rho), _ptr(sn2), _ptr(bias)))

    def capacity(self):
        """Observations the fitted handle can hold before `append` needs a refit."""
        cap = C.c_int()
        self._check(self._lib.bo_fit_capacity(self._h, C.byref(cap)))
        return cap.value

    def append(self, X, y):
        """Incremental refit (bo_append): O(n^2) per new observation."""
        X = f64(X, 2)
        y = f64(y, 1)
        d = getattr(self, "d", 0)
        if d and (X.shape[1] != d or X.shape[0] != len(y)):
            raise ValueError("append: expected (m, %d) inputs and (m,) outputs" % d)
        self._check(self._lib.bo_append(self._h, X.shape[0], _ptr(X), _ptr(y)))
        self.n = getattr(self, "n", 0) + X.shape[0]

    def loglik(self):
        out = np.empty(self.S)
        self._check(self._lib.bo_loglik(self._h, _ptr(out)))
        return out

    def loglik_fit(self, kernel, X, y, ell, rho, sn2, bias):
        """Log marginal likelihood of S hyper-samples from Gram + Cholesky alone (bo_loglik_fit): the fitted state
        of the handle, if any, is not touched."""
        X = f64(X, 2)
        y = f64(y, 1)
        ell = f64(ell, 2)
        rho, sn2, bias = f64(rho), f64(sn2), f64(bias)
        n, d = X.shape
        S = ell.shape[0]
        if ell.shape[1] != d or len(rho) != S or len(sn2) != S or len(bias) != S or len(y) != n:
            raise ValueError("inconsistent fit shapes")
        out = np.empty(S)
        self._check(self._lib.bo_loglik_fit(self._h, KERNEL_IDS[kernel], n, d, S, _ptr(X), _ptr(y), _ptr(ell),
                                            _ptr(rho), _ptr(sn2), _ptr(bias), _ptr(out)))
        return out

    def factor(self, which, s=0):
        code = {"L": 0, "W": 1, "alpha": 2, "beta": 3}[which]
        out = np.empty((self.n, self.n) if code < 2 else (self.n,))
        self._check(self._lib.bo_get_factor(self._h, int(s), code, _ptr(out)))
        return out

    # -- scoring ------------------------------------------------------------------
    def score(self, acq, param, X, grad=False, want_values=True, want_best=False):
        """Host-pointer scoring call.  Returns (values|None, grad|None, best|None)."""
        X = f64(X, 2)
        M, d = X.shape
        if d != self.d:
            raise ValueError("candidates have dimension %d, model has %d" % (d, self.d))
        val = np.empty(M) if want_values else None
        g = np.empty((M, d)) if grad else None
        bv, bi = C.c_double(), C.c_int64()
        self._check(self._lib.bo_score(self._h, int(acq), float(param), M, _ptr(X), PTR_HOST, _ptr(val), _ptr(g),
                                       C.byref(bv) if want_best else None, C.byref(bi) if want_best else None))
        return val, g, ((bv.value, bi.value) if want_best else None)

    def score_device(self, acq, param, M, xc_ptr, val_ptr=None, grad_ptr=None, want_best=True):
        """Device-pointer scoring call (raw addresses, e.g. torch `data_ptr()`)."""
        bv, bi = C.c_double(), C.c_int64()
        self._check(self._lib.bo_score(self._h, int(acq), float(param), int(M), _ptr(xc_ptr), PTR_DEVICE,
                                       _ptr(val_ptr), _ptr(grad_ptr),
                                       C.byref(bv) if want_best else None, C.byref(bi) if want_best else None))
        return (bv.value, bi.value) if want_best else None

    # -- incumbents that stay on the device (cross-rank exchange) -------------------------------
    def score_incumbent(self, acq, param, M, xc, offset=0, val_ptr=None, flags=PTR_DEVICE):
        """Score and leave the (value, index + offset) record on the device; returns its address.
        `xc`: device address (flags=PTR_DEVICE), host array (PTR_HOST) or None (PTR_STAGED)."""
        rec = C.c_void_p()
        if flags == PTR_HOST:
            xc = f64(xc, 2)
        self._check(self._lib.bo_score_incumbent(self._h, int(acq), float(param), int(M), _ptr(xc), int(flags),
                                                 _ptr(val_ptr), int(offset), C.byref(rec)))
        return rec.value

    def thompson_incumbents(self, M, xc, offset=0, flags=PTR_DEVICE):
        """Per-draw arg max left on the device as ndraw records; returns (address, ndraw)."""
        rec, nd = C.c_void_p(), C.c_int()
        if flags == PTR_HOST:
            xc = f64(xc, 2)
        self._check(self._lib.bo_thompson_incumbents(self._h, int(M), _ptr(xc), int(flags), int(offset),
                                                     C.byref(rec), C.byref(nd)))
        return rec.value, nd.value

    def incumbent_merge(self, records_ptr, count, k):
        """(values (k,), indices (k,)) merged from count x k device records."""
        val = np.empty(k)
        idx = np.empty(k, dtype=np.int64)
        self._check(self._lib.bo_incumbent_merge(self._h, _ptr(records_ptr), int(count), int(k), _ptr(val), _ptr(idx)))
        return val, idx

    # -- device-side candidate grid ---------------------------------------------------
    def sobol(self, d, start, M, bounds=None, out="host"):
        """Points [start, start + M) of the unscrambled Sobol sequence in `bounds` ((d, 2), default unit
        cube), generated on the device from SciPy's direction numbers.  out='host' returns the (M, d)
        array; out='staged' leaves the grid in the handle for `score_staged` (no candidate copy at all)."""
        sv, bits = sobol_directions(d)
        lo = hi = None
        if bounds is not None:
            b = f64(bounds, 2)
            lo, hi = np.ascontiguousarray(b[:, 0]), np.ascontiguousarray(b[:, 1])
        res = np.empty((M, d)) if out == "host" else None
        self._check(self._lib.bo_candidates_sobol(self._h, int(d), int(bits), _ptr(sv), _ptr(lo), _ptr(hi),
                                                  int(start), int(M), _ptr(res), PTR_HOST))
        return res

    def score_staged(self, acq, param, M, want_values=False, want_best=True):
        """Score the grid left in the handle by `sobol(..., out='staged')`."""
        val = np.empty(M) if want_values else None
        bv, bi = C.c_double(), C.c_int64()
        self._check(self._lib.bo_score(self._h, int(acq), float(param), int(M), None, PTR_STAGED, _ptr(val), None,
                                       C.byref(bv) if want_best else None, C.byref(bi) if want_best else None))
        return val, ((bv.value, bi.value) if want_best else None)

    def predict(self, X, grad=False):
        X = f64(X, 2)
        M, d = X.shape
        if d != self.d:
            raise ValueError("points have dimension %d, model has %d" % (d, self.d))
        mu, s2 = np.empty(M), np.empty(M)
        dmu = np.empty((M, d)) if grad else None
        ds2 = np.empty((M, d)) if grad else None
        self._check(self._lib.bo_predict(self._h, M, _ptr(X), PTR_HOST, _ptr(mu), _ptr(s2), _ptr(dmu), _ptr(ds2)))
        return (mu, s2, dmu, ds2) if grad else (mu, s2)

    def topk(self, k):
        """(indices, values) of the k best values of the last `score`, descending, lowest index first among ties;
        shorter than k when fewer values are comparable (NaN never ranks)."""
        idx = np.empty(k, dtype=np.int64)
        val = np.empty(k)
        self._check(self._lib.bo_topk(self._h, int(k), _ptr(idx), _ptr(val)))
        keep = idx >= 0
        return idx[keep], val[keep]

    def set_option(self, key, value):
        """Tuning knobs that do not change results (bo_set_option), e.g. ("oz_cluster", 2)."""
        self._check(self._lib.bo_set_option(self._h, key.encode(), float(value)))

    def set_rescue(self, on=True, tol=5e-7, floor_rel=1e-12):
        """FP64 rescue pass of the int8-slice path (bo_set_rescue)."""
        self._check(self._lib.bo_set_rescue(self._h, 1 if on else 0, float(tol), float(floor_rel)))

    def error_bound(self):
        """errk (S,): on the int8 path |s2 - s2_exact| <= errk[s] sqrt((rho_s - s2) rho_s) at the selected level."""
        out = np.empty(self.S)
        self._check(self._lib.bo_ozaki_error_bound(self._h, _ptr(out)))
        return out

    def rescue_info(self):
        """(ran_int8_path, candidates re-scored in FP64, candidates) of the last score / predict call."""
        path, flagged, total = C.c_int(), C.c_int64(), C.c_int64()
        self._check(self._lib.bo_rescue_info(self._h, C.byref(path), C.byref(flagged), C.byref(total)))
        return bool(path.value), flagged.value, total.value

    def tier_info(self):
        """Tiers of the last int8 pass (bo_tier_info): dict with the levels as (slices, extra_group) pairs --
        `first` (candidate chunk 0), `rest` (the other chunks), `tier2` (re-score level of the flagged list, None if
        it went straight to FP64) -- and the counts `first_flagged`, `fp64_rescored`."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        ff, fr = C.c_int64(), C.c_int64()
        self._check(self._lib.bo_tier_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(ff), C.byref(fr)))
        lvl = lambda L: (L >> 1, bool(L & 1)) if L else None
        return dict(first=lvl(a.value), rest=lvl(b.value), tier2=lvl(c.value), first_flagged=ff.value, fp64_rescored=fr.value)

    def set_precision(self, prec, tol=1e-9):
        self._check(self._lib.bo_set_precision(self._h, int(prec), float(tol)))

    def precision_info(self):
        """(path, slices, extra_group) of the scoring contraction."""
        prec, level = C.c_int(), C.c_int()
        self._check(self._lib.bo_precision_info(self._h, C.byref(prec), C.byref(level)))
        return prec.value, level.value // 2, bool(level.value & 1)

    # -- Thompson -------------------------------------------------------------------
    def thompson_set(self, W, b, theta, scale, bias):
        W = f64(W, 3)
        b = f64(b, 2)
        theta = f64(theta, 2)
        scale, bias = f64(scale), f64(bias)
        nW, m, d = W.shape
        ndraw = theta.shape[0]
        if b.shape != (nW, m) or theta.shape[1] != m or len(scale) != ndraw or len(bias) != ndraw:
            raise ValueError("inconsistent Thompson shapes")
        self.th_shape = (ndraw, m, d)
        self._check(self._lib.bo_thompson_set(self._h, ndraw, nW, m, d, _ptr(W), _ptr(b), _ptr(theta),
                                              _ptr(scale), _ptr(bias)))

    def thompson_build(self, X, y, rho, sn2, bias, W, b, noise):
        """Build weight-space posterior draws on the device (bo_thompson_build) and install them for
        `thompson_eval`; returns theta (ndraw, m)."""
        X = f64(X, 2)
        y = f64(y, 1)
        W = f64(W, 3)
        b = f64(b, 2)
        noise = f64(noise, 2)
        n, d = X.shape
        nW, m, dw = W.shape
        ndraw = noise.shape[0]
        if dw != d or b.shape != (nW, m) or noise.shape[1] != m or len(y) != n or nW not in (1, ndraw):
            raise ValueError("inconsistent Thompson build shapes")
        theta = np.empty((ndraw, m))
        self._check(self._lib.bo_thompson_build(self._h, n, d, _ptr(X), _ptr(y), float(rho), float(sn2), float(bias),
                                                ndraw, nW, m, _ptr(W), _ptr(b), _ptr(noise), _ptr(theta)))
        self.th_shape = (ndraw, m, d)
        return theta

    def thompson_eval(self, X, grad=False, want_values=True, want_best=False):
        X = f64(X, 2)
        ndraw, m, d = self.th_shape
        M = X.shape[0]
        if X.shape[1] != d:
            raise ValueError("points have dimension %d, draw has %d" % (X.shape[1], d))
        out = np.empty((ndraw, M)) if want_values else None
        g = np.empty((ndraw, M, d)) if grad else None
        bv = np.empty(ndraw) if want_best else None
        bi = np.empty(ndraw, dtype=np.int64) if want_best else None
        self._check(self._lib.bo_thompson_eval(self._h, M, _ptr(X), PTR_HOST, _ptr(out), _ptr(g), _ptr(bv), _ptr(bi)))
        return out, g, ((bv, bi) if want_best else None)

    def thompson_eval_device(self, M, xc_ptr, out_ptr=None):
        ndraw = self.th_shape[0]
        bv = np.empty(ndraw)
        bi = np.empty(ndraw, dtype=np.int64)
        self._check(self._lib.bo_thompson_eval(self._h, int(M), _ptr(xc_ptr), PTR_DEVICE, _ptr(out_ptr), None,
                                               _ptr(bv), _ptr(bi)))
        return bv, bi

    # -- stand-alone linear algebra ----------------------------------------------------
    def cholesky(self, A):
        A = np.array(A, dtype=np.float64, order="C", copy=True)
        batch = 1 if A.ndim == 2 else A.shape[0]
        n = A.shape[-1]
        info = np.zeros(batch, dtype=np.int32)
        self._check(self._lib.bo_cholesky(self._h, n, batch, _ptr(A), PTR_HOST, _ptr(info)))
        return A

    def cholesky_device(self, n, batch, ptr):
        info = np.zeros(batch, dtype=np.int32)
        self._check(self._lib.bo_cholesky(self._h, int(n), int(batch), _ptr(ptr), PTR_DEVICE, _ptr(info)))
        return info

    def gram(self, kernel, X, ell, rho, sn2):
        X = f64(X, 2)
        ell = f64(ell)
        n, d = X.shape
        K = np.empty((n, n))
        self._check(self._lib.bo_gram(self._h, KERNEL_IDS[kernel], n, d, _ptr(X), _ptr(ell), float(rho),
                                      float(sn2), _ptr(K), PTR_HOST))
        return K

    # -- profiler ---------------------------------------------------------------------------
    def profile(self, on=True):
        self._check(self._lib.bo_profile_enable(self._h, 1 if on else 0))

    def profile_reset(self):
        self._check(self._lib.bo_profile_reset(self._h))

    def profile_report(self):
        cnt = C.c_int()
        self._check(self._lib.bo_profile_count(self._h, C.byref(cnt)))
        out = {}
        buf = C.create_string_buffer(128)
        for i in range(cnt.value):
            launches, ms = C.c_int64(), C.c_double()
            self._check(self._lib.bo_profile_get(self._h, i, buf, 128, C.byref(launches), C.byref(ms)))
            out[buf.value.decode()] = dict(launches=launches.value, total_ms=ms.value)
        return out

    def microbench(self, kind, iters=40000):
        """Measured FP64 roof in TFLOP/s: kind 'dmma' (tensor core) or 'dfma'; or dependent-issue
        latencies in cycles: 'lat_dfma', 'lat_rcp', 'lat_rsqrt', 'lat_syncthreads', 'lat_mbarrier', 'lat_dmma'."""
        kinds = {"dmma": 0, "dfma": 1, "lat_dfma": 2, "lat_rcp": 3, "lat_rsqrt": 4, "lat_syncthreads": 5,
                 "lat_mbarrier": 6, "lat_dmma": 7, "dmma_dfma_mix": 8}
        v = C.c_double()
        self._check(self._lib.bo_microbench(self._h, kinds[kind], int(iters), C.byref(v)))
        return v.value

    def ozaki_debug(self, S, X, want_acc=True, want_slices=True, extra=False):
        """Self-test hook: int8-slice path on hyper-sample 0 for candidates X."""
        X = f64(X, 2)
        mc = X.shape[0]
        npad = -(-self.n // 128) * 128
        mcp = -(-mc // 128) * 128
        mu, s2 = np.empty(mc), np.empty(mc)
        acc = np.empty((npad // 64, S + (1 if extra else 0), 128, 64), dtype=np.int32) if want_acc else None
        ws = np.empty((S, npad, npad), dtype=np.int8) if want_slices else None
        ks = np.empty((S, mcp, npad), dtype=np.int8) if want_slices else None
        rs = np.empty(npad)
        self._check(self._lib.bo_ozaki_debug(self._h, int(S), 1 if extra else 0, mc, _ptr(X), _ptr(mu), _ptr(s2), _ptr(acc),
                                             _ptr(ws), _ptr(ks), _ptr(rs)))
        return dict(mu=mu, s2=s2, acc=acc, ws=ws, ks=ks, rowscale=rs)

    def launch_count(self):
        v = C.c_int64()
        self._check(self._lib.bo_launch_count(self._h, C.byref(v)))
        return v.value
