"""Small end-to-end exercise of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
fit (Gram, Cholesky step kernel with its inter-CTA flags, trtri), FP64 scoring + gradients, the int8-slice
tcgen05/TMA/TMEM path, top-k, device Sobol grid, incremental append, batched stand-alone Cholesky, Thompson on
both paths.  Shapes are small: the sanitizer slows kernels 10-100x."""
import os
import sys

import numpy as np

sys.path.insert(0, '.')
from scipy.stats import qmc
from pybo_b200 import _lib, models

rng = np.random.RandomState(0)
n, d, M = 300, 3, 1500
X = rng.rand(n, d)
y = np.sin(X.sum(1)) + 0.01 * rng.randn(n)
rho, bias = float(np.ptp(y)), float(y.mean())
ctx = _lib.Context(0)
ctx.fit("se", X[:-2], y[:-2], 0.3 * np.ones((2, d)), [rho, 1.1 * rho], [1e-4, 2e-4], [bias, bias])
ctx.append(X[-2:], y[-2:])
Xc = qmc.Sobol(d=d, scramble=False).random_base2(11)[:M]
mu, s2 = ctx.predict(Xc)
target = float(ctx.predict(X)[0].max())
v0, _, b0 = ctx.score(1, target, Xc, want_best=True)
vg, g, _ = ctx.score(1, target, Xc[:5], grad=True)
vg, g, _ = ctx.score(2, target, Xc[:100], grad=True)
idx, val = ctx.topk(10)
ctx.set_precision(1, 1e-8)
v8, _, b8 = ctx.score(1, target, Xc, want_best=True)
ctx.set_precision(1, 4.5)
v8b, _, _ = ctx.score(3, 2.0, Xc, want_best=True)
ctx.set_precision(0)
ctx.sobol(d, 0, 1024, out="staged")
ctx.score_staged(1, target, 1024)
print("loglik", ctx.loglik(), ctx.loglik_fit("se", X, y, 0.3 * np.ones((2, d)), [rho, 1.1 * rho], [1e-4, 2e-4], [bias, bias]))
# rescue pass of the int8 path (coarse level so that candidates are flagged), large-list compaction, incumbent records
ctx.set_precision(1, 4.0)
vr, _, br = ctx.score(1, target, Xc, want_best=True)
print("rescue", ctx.rescue_info())
Xbig = qmc.Sobol(d=d, scramble=False).random_base2(14)
ctx.set_rescue(True, 1e-12, 1e-12)
ctx.score(2, target, Xbig, want_best=True)
print("rescue (everything flagged)", ctx.rescue_info())
ctx.set_rescue(True)
# tiers of the int8 path (sanitize.sh shrinks the chunks with BO_OZ_CHUNK_TILES so that 2^14 candidates are >= 8 chunks):
# pilot passes -> whole pass one half-level down; pilot fails -> mixed levels; flagged list one tier up, then FP64
ctx.set_precision(1, 1e-8)
ctx.set_option("oz_tier_min", 1)
for frac in (1.0, 0.0):
    ctx.set_option("oz_tier_frac", frac)
    ctx.set_rescue(True)
    vt, _, bt = ctx.score(1, target, Xbig, want_best=True)
    print("tiers", ctx.tier_info(), bt)
    mt, st = ctx.predict(Xbig)
    print("tiers (predict)", ctx.tier_info())
ctx.set_option("oz_tier_frac", 0.1)
ctx.set_option("oz_tier_min", 4096)
if os.environ.get("SAN_ONLY") == "tiers":
    print("ok (tiers only)")
    sys.exit(0)
ctx.set_precision(0)
rec = ctx.score_incumbent(1, target, len(Xc), Xc, offset=5, flags=0)
print("incumbent", ctx.incumbent_merge(rec, 1, 1))
A1 = rng.randn(600, 600)
A1 = A1 @ A1.T + 600 * np.eye(600)
L1 = ctx.cholesky(A1)                       # one matrix, 10 blocks: the dataflow kernel when BO_CHOL_FLOW_MIN <= 10
print("chol (single) err", np.abs(L1 - np.linalg.cholesky(A1)).max())
A = rng.randn(2, 200, 200)
A = A @ A.transpose(0, 2, 1) + 200 * np.eye(200)
L = ctx.cholesky(A)
print("chol err", np.abs(L - np.linalg.cholesky(A)).max())
K = ctx.gram("matern52", X, 0.3 * np.ones(d), rho, 1e-4)
gp = models.make_gp(1e-4, rho, 0.3 * np.ones(d), bias)
gp.add_data(X, y)
tb = models.ThompsonBatch(gp, m=128, ndraw=64, rng=1)
F = tb.get(Xc)
tb.set_precision("int8", 1e-8)
F8 = tb.get(Xc)
bv, bi = tb.argmax(Xc)
tp = models.ThompsonBatch(gp, m=40, ndraw=6, rng=3, shared_basis=False)      # one feature system per draw, built on the device
print("per-draw bases", tp.argmax(Xc)[1])
fs = gp.sample_f(50, rng=2)
fv, fg = fs.get(Xc[:7], grad=True)
print("ok", b0, b8, float(np.abs(v8 - v0).max()), float(np.abs(F8 - F).max()))
