"""Calibration of the int8-slice path's error bound (GPU): for several shapes and slice levels, the
observed error of s2 = rho - |W k*|^2 against the FP64 path, next to the a-priori model
est(level) = 8 sqrt(np) 2^emax sqrt(rho) 256^-S (/32 with the extra pair group) that bo_ozaki_choose_slices
uses.  The rescue pass flags a candidate when  K * est * sqrt(q rho)  (q = rho - s2) is too large for the
acquisition's tolerance; this tool measures  max |ds2| / (est sqrt(q rho))  so K can be chosen with margin.
Writes one JSON line per (shape, level)."""
import json
import sys

import numpy as np

sys.path.insert(0, '.')
from scipy.stats import qmc
from pybo_b200 import _lib

SHAPES = [("se", 4096, 8, 0.25, 1e-6, 0), ("se", 1024, 4, 0.25, 1e-6, 0), ("matern52", 4096, 8, 0.25, 1e-6, 3),
          ("se", 2048, 8, 0.25, 1e-6, 0), ("se", 640, 3, 0.3, 1e-4, 5), ("matern52", 383, 2, 0.3, 1e-4, 6),
          ("se", 1025, 12, 0.5, 1e-4, 7), ("se", 200, 1, 0.3, 1e-4, 8), ("se", 4096, 16, 0.5, 1e-4, 4)]
LEVELS = (3.5, 4, 4.5, 5, 5.5, 6)


def main():
    ctx = _lib.Context(0)
    for kernel, n, d, ellv, sn2, seed in SHAPES:
        rng = np.random.RandomState(seed)
        X = rng.rand(n, d)
        y = np.sin(X.sum(1)) + 0.01 * rng.randn(n)
        rho, bias = float(y.max() - y.min()), float(y.mean())
        ctx.set_precision(0)
        ctx.fit(kernel, X, y, ellv * np.ones((1, d)), [rho], [sn2], [bias])
        W = ctx.factor("W")
        rowmax = np.abs(W).max(axis=1)
        e = np.frexp(rowmax)[1]
        emax = int(e.max())
        npad = -(-n // 128) * 128
        del W
        Xc = np.concatenate([qmc.Sobol(d=d, scramble=False).random_base2(16),
                             np.clip(X[rng.randint(0, n, 4096)] + 1e-3 * rng.randn(4096, d), 0, 1),
                             np.clip(X[rng.randint(0, n, 4096)] + 3e-2 * rng.randn(4096, d), 0, 1)])
        mu0, s20 = ctx.predict(Xc)
        q = np.maximum(rho - s20, 0.0)
        for lev in LEVELS:
            S, extra = int(lev), (lev - int(lev)) >= 0.5
            if npad * S >= (1 << 17):
                continue
            ctx.set_precision(1, float(lev))
            mu, s2 = ctx.predict(Xc)
            est = 8.0 * np.sqrt(npad) * 2.0 ** emax * np.sqrt(rho) * 2.0 ** (-8 * S) / (32.0 if extra else 1.0)
            ds2 = np.abs(s2 - s20)
            ratio = ds2 / np.maximum(est * np.sqrt(q * rho), 1e-300)
            # row-resolved model: sigma_i ~ 2^e_i; D = sqrt(sum_i 4^(e_i - emax)) / sqrt(np)
            rowfac = float(np.sqrt(np.sum(4.0 ** (e - emax)) / npad))
            print(json.dumps(dict(kernel=kernel, n=n, d=d, level=lev, emax=emax, est=est, rowfac=rowfac,
                                  max_ds2_over_rho=float(ds2.max() / rho), rms_ds2_over_rho=float(np.sqrt(np.mean(ds2 ** 2)) / rho),
                                  max_ratio=float(ratio.max()), p999_ratio=float(np.percentile(ratio, 99.9)),
                                  median_ratio=float(np.median(ratio)), max_dmu=float(np.abs(mu - mu0).max()),
                                  s2min_over_rho=float(s20.min() / rho), argmax_ratio_q=float(q[np.argmax(ratio)] / rho))),
                  flush=True)
        ctx.set_precision(0)


if __name__ == "__main__":
    main()
