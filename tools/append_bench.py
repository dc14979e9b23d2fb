"""bo_append timing at n = 4096 (GPU): wall clock per call and the two triangular matrix-vector kernels (CUDA events)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from pybo_b200 import _lib
n, d, k = 4096, 8, 32
rng = np.random.RandomState(0)
X = rng.rand(n, d); y = np.sin(X.sum(1)) + 0.01 * rng.randn(n)
c = _lib.Context(0)
c.fit("se", X[:n - k], y[:n - k], 0.25 * np.ones((1, d)), [2.0], [1e-6], [0.0])
c.append(X[n - k], y[n - k:n - k + 1]); c.sync()
t0 = time.perf_counter()
for i in range(n - k + 1, n - 8):
    c.append(X[i], y[i:i + 1])
c.sync()
dt = (time.perf_counter() - t0) / (k - 9)
c.profile(True); c.profile_reset()
for i in range(n - 8, n):
    c.append(X[i], y[i:i + 1])
c.sync()
prof = c.profile_report()
print("ms per append %.4f  (%.2f TB/s on n^2 * 8 B)" % (dt * 1e3, n * n * 8 / dt / 1e12))
for kname, v in prof.items():
    print("   %-22s %.1f us" % (kname, 1e3 * v["total_ms"] / v["launches"]))
