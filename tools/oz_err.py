"""Error statistics of the int8-slice path vs the FP64 path at the headline shape (GPU)."""
import sys, numpy as np
sys.path.insert(0, '.')
from scipy.stats import qmc
from pybo_b200 import _lib
import os
n, d, M = 4096, 8, int(os.environ.get("OZ_ERR_M", "40000"))
rng = np.random.RandomState(0)
X = rng.rand(n, d); y = np.sin(X.sum(1)) + 0.01 * rng.randn(n)
rho, bias = float(y.max() - y.min()), float(y.mean())
ctx = _lib.Context(0)
ctx.fit("se", X, y, 0.25 * np.ones((1, d)), [rho], [1e-6], [bias])
Xc = qmc.Sobol(d=d, scramble=False).random_base2(20)[:M]
target = float(ctx.predict(X)[0].max())
mu0, s20 = ctx.predict(Xc)
ei0, _, _ = ctx.score(1, target, Xc)
z = (mu0 - target) / np.sqrt(s20)
print("z range", z.min(), z.max(), "s2/rho range", s20.min() / rho, s20.max() / rho, "EI max", ei0.max())
import time
for S in (3.5, 4, 4.5, 5, 5.5, 6):
    ctx.set_precision(1, float(S))
    mu, s2 = ctx.predict(Xc)
    ctx.sync(); t0 = time.perf_counter()
    ei, _, _ = ctx.score(1, target, Xc)
    dt = time.perf_counter() - t0
    floor = 1e-9 * ei0.max()
    rel = np.abs(ei - ei0) / np.maximum(np.abs(ei0), floor)
    i = int(np.argmax(rel))
    print("S=%s (%.0f ms incl. copies) argmax same=%s  max|dmu|=%.2e mean(dmu)=%.2e  max|ds2|/rho=%.2e mean(ds2)/rho=%.2e | EI rel: max=%.2e p99.9=%.2e median=%.2e | worst: z=%.2f ei0=%.3e"
          % (S, dt * 1e3, int(np.argmax(ei)) == int(np.argmax(ei0)), np.abs(mu - mu0).max(), (mu - mu0).mean(), np.abs(s2 - s20).max() / rho, (s2 - s20).mean() / rho,
             rel.max(), np.percentile(rel, 99.9), np.median(rel), z[i], ei0[i]))
