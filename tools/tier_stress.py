"""Tiered levels of the int8 path at real scale (GPU): random shapes n = 1500..4096, d = 2..16, both kernels, EI / PI / UCB,
2^19 Sobol candidates (16 chunks of 32 768, so the pilot chunk and the lowered main level engage with the production
chunk size).  Each case: int8 path with the tiers on vs the FP64 path over all candidates (1e-6, floor 1e-12 max|ref|,
identical arg max) and vs the oracle on a random subset of 2048."""
import sys

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from scipy.stats import qmc
from conftest import rel_err
from oracle import GPOracle, ucb_beta, ucb_index
from pybo_b200 import _lib

bad = []
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    rng = np.random.RandomState(500 + seed)
    n = int(rng.choice([1500, 2048, 3000, 4096]))
    d = int(rng.choice([2, 4, 8, 12, 16]))
    kernel = str(rng.choice(["se", "matern52"]))
    acq = int(rng.choice([1, 2, 3]))
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    ell = 0.25 * max(1.0, np.sqrt(d / 8.0)) * np.exp(0.1 * rng.randn(d))
    rho, sn2, bias = float(np.ptp(y)), float(10 ** rng.uniform(-6, -3)), float(y.mean())
    ctx = _lib.Context(0)
    ctx.fit(kernel, X, y, ell[None], [rho], [sn2], [bias])
    Xc = qmc.Sobol(d=d, scramble=False).random_base2(19)
    gp = GPOracle(sn2, rho, ell, bias, kernel)
    gp.add_data(X, y)
    target = float(ctx.predict(X)[0].max())
    param = float(ucb_beta(n)) if acq == 3 else (target if acq == 1 else target + 0.05)
    ctx.set_precision(0, 1e-8)
    ref, _, rbest = ctx.score(acq, param, Xc, want_best=True)
    ctx.set_precision(1, 1e-8)
    val, _, best = ctx.score(acq, param, Xc, want_best=True)
    ran8 = ctx.rescue_info()[0]
    t = ctx.tier_info()
    sub = rng.choice(len(Xc), 2048, replace=False)
    mu, s2 = gp.predict(Xc[sub])
    oref = gp.get_improvement(param, Xc[sub]) if acq == 1 else (gp.get_tail(param, Xc[sub]) if acq == 2 else ucb_index(param, mu, s2))
    floor = 1e-12 if d > 2 else 1e-9
    scale = float(np.abs(ref).max())
    e_gpu = rel_err(val, ref, floor)
    e_orc = float(np.max(np.abs(val[sub] - oref) / np.maximum(np.abs(oref), max(floor, 1e-9) * scale)))
    ok = e_gpu < 1e-6 and e_orc < 1e-6 and best[1] == rbest[1]
    print("seed %2d n=%d d=%2d %-8s acq=%d sn2=%.1e  int8=%s tiers=%s  vs fp64 %.2e  vs oracle %.2e  argmax %s  %s"
          % (seed, n, d, kernel, acq, sn2, ran8, t, e_gpu, e_orc, best[1] == rbest[1], "ok" if ok else "FAIL"), flush=True)
    if not ok:
        bad.append(seed)
    ctx.close()
print("failures:", bad)
