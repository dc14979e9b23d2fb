import sys, numpy as np, torch
sys.path.insert(0, '.')
from pybo_b200 import _lib
ctx = _lib.Context(0)
rng = np.random.RandomState(0)
B = rng.randn(64, 80); A = torch.from_numpy(B @ B.T + np.eye(64)).cuda()
for batch in (1, 32):
    Ab = A.repeat(batch, 1, 1).contiguous(); work = torch.empty_like(Ab)
    ctx.profile(True); ctx.profile_reset()
    for _ in range(50):
        work.copy_(Ab); torch.cuda.synchronize()
        try: ctx.cholesky_device(64, batch, work.data_ptr())
        except Exception: pass
    r = ctx.profile_report()["potrf64_kernel"]
    print("batch", batch, "potrf64 avg us", 1e3 * r["total_ms"] / r["launches"])
