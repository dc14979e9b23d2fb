"""Fit cost (Gram + Cholesky + W = L^-1 + alpha/beta/logdet), batched over hyper-samples (GPU)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from pybo_b200 import _lib
ctx = _lib.Context(0)
for (n, d, S) in ((4096, 8, 1), (2048, 8, 32), (1024, 4, 1), (256, 2, 10)):
    rng = np.random.RandomState(0)
    X = rng.rand(n, d); y = np.sin(X.sum(1)) + 0.01 * rng.randn(n)
    ell = 0.25 * np.ones((S, d)) * np.exp(0.1 * rng.randn(S, d)); rho = np.full(S, 2.0); sn2 = np.full(S, 1e-6); bias = np.zeros(S)
    ctx.fit("se", X, y, ell, rho, sn2, bias); ctx.sync()
    ctx.profile(True); ctx.profile_reset()
    t0 = time.perf_counter(); ctx.fit("se", X, y, ell, rho, sn2, bias); ctx.sync(); dt = time.perf_counter() - t0
    rep = ctx.profile_report(); ctx.profile(False)
    top = sorted(rep.items(), key=lambda kv: -kv[1]["total_ms"])[:5]
    print("n=%d d=%d S=%d: fit %.2f ms (%.1f TFLOP/s on 2/3 n^3 S) | %s" % (n, d, S, dt * 1e3, 2 / 3 * n ** 3 * S / dt / 1e12,
          ", ".join("%s %.2f" % (k.replace("_kernel", ""), v["total_ms"]) for k, v in top)))
