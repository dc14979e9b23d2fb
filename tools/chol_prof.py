"""Per-kernel breakdown of the stand-alone Cholesky (GPU)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pybo_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.RandomState(0)
X = rng.rand(n, 8)
ctx = _lib.Context(0)
K = torch.from_numpy(ctx.gram("se", X, 0.25 * np.ones(8), 2.0, 1e-6)).cuda()
work = torch.empty_like(K)
for rep in range(3):
    work.copy_(K); torch.cuda.synchronize()
    t0 = time.perf_counter(); ctx.cholesky_device(n, 1, work.data_ptr()); dt = time.perf_counter() - t0
    print("wall %.3f ms" % (dt * 1e3))
ctx.profile(True); ctx.profile_reset()
work.copy_(K); torch.cuda.synchronize()
t0 = time.perf_counter(); ctx.cholesky_device(n, 1, work.data_ptr()); dt = time.perf_counter() - t0
rep = ctx.profile_report()
print("profiled wall %.3f ms" % (dt * 1e3))
for k, v in rep.items():
    print("  %-22s launches %4d  total %.3f ms  avg %.1f us" % (k, v["launches"], v["total_ms"], 1e3 * v["total_ms"] / max(1, v["launches"])))
L = work.cpu().numpy(); ref = np.linalg.cholesky(K.cpu().numpy())
print("max err", np.abs(L - ref).max())
