"""Latency of the L-BFGS callback shape f(x[None], grad=True) (reference solvers/lbfgs.py:56-58)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
from pybo_b200 import _lib, models, policies, solvers
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 8
rng = np.random.RandomState(0)
X = rng.rand(n, d); y = np.sin(X.sum(1)) + 0.01 * rng.randn(n)
gp = models.make_gp(1e-6, float(y.max() - y.min()), 0.25 * np.ones(d), float(y.mean()))
gp.add_data(X, y)
t0 = time.perf_counter(); ctx = gp._ensure_fit(); ctx.sync(); print("fit %.1f ms" % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter(); ctx.fit("se", X, y, 0.25 * np.ones((1, d)), [2.0], [1e-6], [0.0]); ctx.sync(); print("refit %.1f ms" % (1e3 * (time.perf_counter() - t0)))
index = policies.EI(gp, None, list(X[:50]))
for M in (1, 10, 128):
    x = rng.rand(M, d)
    for _ in range(3): index(x, grad=True)
    t0 = time.perf_counter()
    for _ in range(20): index(x, grad=True)
    print("M=%d grad call: %.3f ms" % (M, 1e3 * (time.perf_counter() - t0) / 20))
    t0 = time.perf_counter()
    for _ in range(20): index(x)
    print("M=%d value call: %.3f ms" % (M, 1e3 * (time.perf_counter() - t0) / 20))
bounds = np.array([[0, 1.0]] * d)
t0 = time.perf_counter(); xb, fb = solvers.solve_lbfgs(index, bounds, ngrid=100000, rng=0); dt = time.perf_counter() - t0
print("solve_lbfgs(ngrid=1e5, nbest=10): %.1f ms, f=%.6f" % (1e3 * dt, fb))
t0 = time.perf_counter(); xc, fc = solvers.solve_lbfgs_batched(index, bounds, ngrid=100000, rng=0); dt = time.perf_counter() - t0
print("solve_lbfgs_batched(ngrid=1e5, nbest=10): %.1f ms, f=%.6f" % (1e3 * dt, fc))
grid = rng.rand(100000, d)
t0 = time.perf_counter(); starts, _ = index.best_of(grid, 10); t1 = time.perf_counter()
xr, fr, nev = solvers.batched_lbfgs(index, grid[starts], bounds); t2 = time.perf_counter()
print("grid scoring %.1f ms; batched refinement of 10 starts %.1f ms (%d batched calls), best %.6f" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), nev, fr.max()))
