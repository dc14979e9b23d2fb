"""BASELINE config 4 shape: Thompson, n=4096 d=16, 256 draws x 2^20 candidates, shared basis (1 GPU)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from pybo_b200 import models
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, d, ndraw, M = 4096, 16, 256, 1 << 20
rng = np.random.RandomState(0)
X = rng.rand(n, d); y = np.sin(X.sum(1)) + 0.01 * rng.randn(n)
gp = models.make_gp(1e-6, float(y.max() - y.min()), 0.25 * np.ones(d), float(y.mean()))
gp.add_data(X, y)
t0 = time.perf_counter(); tb = models.ThompsonBatch(gp, m=m, ndraw=ndraw, rng=0); print("draw construction %.2f s" % (time.perf_counter() - t0))
ctx = tb._context()
xc = torch.quasirandom.SobolEngine(d, scramble=False).draw(M, dtype=torch.float64).cuda()
for _ in range(2): ctx.thompson_eval_device(M, xc.data_ptr())
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): bv, bi = ctx.thompson_eval_device(M, xc.data_ptr())
dt = (time.perf_counter() - t0) / 3
print("m=%d: %.1f ms per pass, %.3e draw-evals/s, %.1f TFLOP/s fp64" % (m, dt * 1e3, ndraw * M / dt, 2.0 * ndraw * M * m / dt / 1e12))
F64 = None
if M <= (1 << 20):
    small = xc[: 1 << 14].cpu().numpy()
    F64 = tb.get(small)
tb.set_precision("int8", 1e-8)
ctx = tb._context()          # (the precision path is written into the handle here)
for _ in range(2): ctx.thompson_eval_device(M, xc.data_ptr())
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): bv8, bi8 = ctx.thompson_eval_device(M, xc.data_ptr())
dt = (time.perf_counter() - t0) / 3
print("int8 path: %.1f ms per pass, %.3e draw-evals/s (%.1f algorithmic TFLOP/s); arg max identical for %d of %d draws"
      % (dt * 1e3, ndraw * M / dt, 2.0 * ndraw * M * m / dt / 1e12, int(np.sum(bi8 == bi)), ndraw))
F8 = tb.get(small)
print("max |F8 - F64| / max|F64| on 2^14 candidates: %.2e" % (np.abs(F8 - F64).max() / np.abs(F64).max()))
ctx.profile(True); ctx.profile_reset(); ctx.thompson_eval_device(M, xc.data_ptr()); ctx.sync()
for k, v in ctx.profile_report().items(): print("  %-26s launches %4d total %.2f ms" % (k, v["launches"], v["total_ms"]))
