"""How many 64x64 blocks of W = L^-1 have an all-zero top (and second) int8 slice? (headline shape)"""
import sys, numpy as np
sys.path.insert(0, '.')
from pybo_b200 import _lib
n, d = 4096, 8
rng = np.random.RandomState(0)
X = rng.rand(n, d); y = np.sin(X.sum(1)) + 0.01 * rng.randn(n)
rho, bias = float(y.max() - y.min()), float(y.mean())
ctx = _lib.Context(0)
def stats(X, y, tag):
    ctx.fit("se", X, y, 0.25 * np.ones((1, d)), [rho], [1e-6], [bias])
    W = ctx.factor("W")
    mx = np.abs(W).max(axis=1); _, e = np.frexp(mx)
    Wn = np.abs(W) * np.exp2(-e)[:, None]                       # < 1
    B = Wn.reshape(n // 64, 64, n // 64, 64).max(axis=(1, 3))   # block maxima
    low = np.tril(np.ones_like(B, dtype=bool), -1)
    t = np.floor(-np.log2(np.maximum(B, 1e-300)) / 7.0)          # leading all-zero slices (approx: < 2^-7 -> top slice rounds to 0 if < 2^-7/2)
    for name, mask in (("strictly lower blocks", low), ("diagonal blocks", np.eye(len(B), dtype=bool))):
        tt = t[mask]
        print(tag, name, "tmin=0: %.3f  tmin=1: %.3f  tmin>=2: %.3f" % ((tt < 1).mean(), ((tt >= 1) & (tt < 2)).mean(), (tt >= 2).mean()))
stats(X, y, "random order |")
# spatial ordering: recursive coordinate bisection (kd-order)
def kd_order(idx, depth=0):
    if len(idx) <= 64: return list(idx)
    k = depth % d; o = idx[np.argsort(X[idx, k])]; h = len(o) // 2
    return kd_order(o[:h], depth + 1) + kd_order(o[h:], depth + 1)
perm = np.array(kd_order(np.arange(n)))
stats(X[perm], y[perm], "kd order     |")
