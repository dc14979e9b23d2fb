"""Cholesky timing (GPU): n = 4096 / 2048 / 1024 single and 32 x 2048 batched, wall clock per call (best of 5) and the
per-kernel breakdown from the library's event profiler."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import torch
from pybo_b200 import _lib

ctx = _lib.Context(0)
roof = max(ctx.microbench("dmma"), ctx.microbench("dfma"))
rng = np.random.RandomState(0)
for n, batch in ((4096, 1), (2048, 1), (1024, 1), (2048, 32), (256, 10)):
    X = rng.rand(n, 8)
    K1 = torch.from_numpy(ctx.gram("se", X, 0.25 * np.ones(8), 1.0, 1e-6)).cuda()
    K = K1.unsqueeze(0).repeat(batch, 1, 1).contiguous()
    work = torch.empty_like(K)
    best = 1e9
    for rep in range(5):
        work.copy_(K)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.cholesky_device(n, batch, work.data_ptr())
        best = min(best, time.perf_counter() - t0)
    L = work[0].cpu().numpy()
    ref = np.linalg.cholesky(K1.cpu().numpy())
    err = np.abs(np.tril(L) - ref).max()
    ctx.profile(True)
    ctx.profile_reset()
    work.copy_(K)
    torch.cuda.synchronize()
    ctx.cholesky_device(n, batch, work.data_ptr())
    prof = ctx.profile_report()
    ctx.profile(False)
    tf = batch * n ** 3 / 3.0 / best / 1e12
    print("n=%d batch=%d: %.3f ms  %.2f TFLOP/s = %.3f of FP64 roof (%.1f)  max|L - lapack| = %.2e" % (n, batch, best * 1e3, tf, tf / roof, roof, err))
    for k, v in sorted(prof.items()):
        print("    %-22s launches %4d  total %.3f ms  avg %.1f us" % (k, v["launches"], v["total_ms"], 1e3 * v["total_ms"] / max(1, v["launches"])))
