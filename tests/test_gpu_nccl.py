"""Two ranks, two GPUs, NCCL: the incumbent exchange that follows the scoring kernels on the library's own stream
(`dist.exchange_incumbents`: packed device records -> one all-gather -> device merge), for the scoring pass and for
the Thompson batch.  Skipped on boxes with fewer than two devices (run with `gpurun --gpus 2`)."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %r)
from scipy.stats import qmc
from pybo_b200 import _lib, dist as bdist, models

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.RandomState(0)
n, d, M = 500, 4, 1 << 15
X = rng.rand(n, d)
y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
rho, bias = float(np.ptp(y)), float(y.mean())
ctx = _lib.Context(local)
ctx.fit("se", X, y, 0.3 * np.ones((1, d)), [rho], [1e-4], [bias])
Xc = qmc.Sobol(d=d, scramble=False).random_base2(15)
Xc[[100, 20000, 30000]] = Xc[100]                          # a three-way tie across both shards if it is the maximum
target = float(ctx.predict(X)[0].max())
full, _, best = ctx.score(1, target, Xc, want_best=True)   # every rank also scores everything: the expected answer
lo, hi = bdist.shard_range(M, rank, world)
xs = torch.from_numpy(np.ascontiguousarray(Xc[lo:hi])).cuda()
for prec in (0, 1):
    ctx.set_precision(prec, 1e-8)
    rec = ctx.score_incumbent(1, target, hi - lo, xs.data_ptr(), offset=lo)
    v, i = bdist.exchange_incumbents(ctx, rec, 1)
    assert int(i[0]) == best[1], (rank, prec, int(i[0]), best[1])
    assert abs(v[0] - best[0]) <= 1e-9 * abs(best[0]), (rank, prec)
ctx.set_precision(0)
# host-staged form gives the same answer
bv, bi = ctx.score_device(1, target, hi - lo, xs.data_ptr(), want_best=True)
hv, hi_ = bdist.reduce_incumbent(bv, bi + lo)
assert hi_ == best[1]
# Thompson batch: 64 draws, per-draw arg max over both shards
gp = models.make_gp(1e-4, rho, 0.3 * np.ones(d), bias, device=local)
gp.add_data(X, y)
tb = models.ThompsonBatch(gp, m=128, ndraw=64, rng=5)
ev, ei = tb.argmax(Xc)
for path in ("fp64", "int8"):
    tb.set_precision(path, 1e-8)
    st = bdist.ShardedThompson(tb)
    gv, gi = st.argmax_device(hi - lo, xs.data_ptr(), lo)
    assert np.array_equal(gi, ei), (rank, path)
    gv2, gi2 = st.argmax(Xc)
    assert np.array_equal(gi2, ei), (rank, path)
dist.barrier()
dist.destroy_process_group()
print("rank %%d ok" %% rank)
''' % ROOT


def test_two_rank_nccl_incumbent_exchange(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "rank 0 ok" in res.stdout and "rank 1 ok" in res.stdout
