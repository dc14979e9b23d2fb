"""Thompson draws built on the device (bo_thompson_build: features on the observations, Phi^T Phi + sn2 I, its
Cholesky factor, the solves) against the oracle's SciPy/LAPACK construction on the same NumPy random stream
(reference policies/simple.py:44-48, BASELINE config 4), for a shared basis and for one basis per draw."""

import numpy as np
import pytest
from scipy.stats import qmc

from conftest import rel_err
from oracle import GPOracle, FourierSampleOracle, thompson_batch_oracle

pytestmark = pytest.mark.gpu


def synth(n, d, kernel="se", seed=0, sn2=1e-3, ellv=0.3):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    gp = GPOracle(sn2, float(np.ptp(y)), ellv * np.ones(d), float(y.mean()), kernel)
    gp.add_data(X, y)
    return gp


def _mine(gp):
    from pybo_b200 import models
    m = models.GP(gp.sn2, gp.rho, gp.ell, gp.bias, kernel=gp.kernel)
    m.add_data(gp.X, gp.Y)
    return m


def _values(W, b, theta, scale, bias, Xc):
    """(ndraw, M) values of the draws, NumPy."""
    if W.shape[0] == 1:
        return (bias + (scale * np.cos(Xc @ W[0].T + b[0])) @ theta.T).T
    return np.array([bias + (scale * np.cos(Xc @ W[r].T + b[r])) @ theta[r] for r in range(len(theta))])


@pytest.mark.parametrize("kernel,n,d,m,ndraw,shared", [("se", 60, 3, 100, 1, True), ("matern52", 300, 5, 200, 16, True),
                                                       ("se", 1000, 16, 512, 64, True), ("se", 150, 4, 100, 12, False),
                                                       ("matern52", 500, 2, 33, 5, False)])
def test_device_built_draws_match_oracle(kernel, n, d, m, ndraw, shared):
    from pybo_b200 import models
    gp = synth(n, d, kernel, seed=n)
    W, b, theta, scale = thompson_batch_oracle(gp, m, ndraw, rng=11, shared_basis=shared)
    tb = models.ThompsonBatch(_mine(gp), m=m, ndraw=ndraw, rng=11, shared_basis=shared)
    assert np.array_equal(tb.W, W) and np.array_equal(tb.b, b)          # same random stream, same order
    Xc = qmc.Sobol(d=d, scramble=False).random_base2(11)
    ref = _values(W, b, theta, scale, gp.bias, Xc)
    F = tb.get(Xc)
    span = np.max(np.abs(ref - gp.bias)) + 1e-300
    # theta solves an ill-conditioned m x m system (cond ~ n rho / (m sn2)); the draws as FUNCTIONS agree to 1e-6 of
    # their own span, which is what the arg max sees
    assert np.max(np.abs(F - ref)) < 1e-6 * span, np.max(np.abs(F - ref)) / span
    bv, bi = tb.argmax(Xc)
    assert np.array_equal(bi, np.argmax(ref, axis=1))
    # device-built theta reproduces the posterior mean of the weights: Phi theta_mean ~ y at the data (shared basis, many draws)
    if shared and ndraw >= 16 and m >= 200:
        fX = tb.get(gp.X[:200])
        assert np.mean((fX.mean(axis=0) - gp.Y[:200]) ** 2) < 0.5 * np.var(gp.Y)


def test_single_draw_is_the_reference_sample_f():
    """`model.sample_f(n, rng).get` (simple.py:48): one draw, same stream as the oracle's FourierSampleOracle."""
    gp = synth(80, 2, "se", seed=5)
    ref = FourierSampleOracle(gp, 100, rng=3)
    draw = _mine(gp).sample_f(100, rng=3)
    Xc = qmc.Sobol(d=2, scramble=False).random_base2(10)
    F, G = draw.get(Xc, grad=True)
    RF, RG = ref.get(Xc, grad=True)
    span = np.max(np.abs(RF - gp.bias))
    assert np.max(np.abs(F - RF)) < 1e-6 * span and np.max(np.abs(G - RG)) < 1e-5 * np.max(np.abs(RG))
    assert draw.argmax(Xc)[1] == int(np.argmax(RF))


def test_sharded_thompson_single_process_equals_unsharded():
    """`ShardedThompson` with world size 1 (no process group): records packed on the device, merged on the device."""
    from pybo_b200 import dist as bdist, models
    gp = synth(200, 4, "se", seed=9)
    tb = models.ThompsonBatch(_mine(gp), m=128, ndraw=32, rng=2)
    Xc = qmc.Sobol(d=4, scramble=False).random_base2(13)
    bv, bi = tb.argmax(Xc)
    halves = []
    for rank in range(2):                                                # emulate two ranks' blocks, merge on the host
        st = bdist.ShardedThompson(tb, rank=rank, world=2)
        lo, hi = bdist.shard_range(len(Xc), rank, 2)
        ctx = tb._context()
        rec, nd = ctx.thompson_incumbents(hi - lo, np.ascontiguousarray(Xc[lo:hi]), offset=lo, flags=0)
        halves.append(ctx.incumbent_merge(rec, 1, nd))
    v = np.array([h[0] for h in halves])
    ix = np.array([h[1] for h in halves])
    best = v.max(axis=0)
    gi = np.where(v == best, ix, np.iinfo(np.int64).max).min(axis=0)
    assert np.array_equal(gi, bi) and np.array_equal(best, bv)
    st = bdist.ShardedThompson(tb, rank=0, world=1)
    v1, i1 = st.argmax(Xc)
    assert np.array_equal(i1, bi) and np.array_equal(v1, bv)


def test_score_incumbent_record_and_merge(ctx):
    rng = np.random.RandomState(1)
    X = rng.rand(120, 3)
    y = np.sin(X.sum(axis=1))
    ctx.fit("se", X, y, 0.3 * np.ones((1, 3)), [1.0], [1e-4], [0.0])
    Xc = qmc.Sobol(d=3, scramble=False).random_base2(12)
    val, _, best = ctx.score(1, 0.5, Xc, want_best=True)
    rec = ctx.score_incumbent(1, 0.5, len(Xc), Xc, offset=1000, flags=0)
    v, i = ctx.incumbent_merge(rec, 1, 1)
    assert v[0] == best[0] and i[0] == best[1] + 1000
