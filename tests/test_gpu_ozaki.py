"""GPU tests of the int8-slice (Ozaki) scoring path on tcgen05: exact integer checks of
the slice planes and the TMEM group accumulators, then parity of mu / s2 / EI against the
FP64 path and the oracle at the slice counts the error model selects."""

import numpy as np
import pytest
from scipy.stats import qmc

from conftest import rel_err
from oracle import GPOracle, kernel_matrix

pytestmark = pytest.mark.gpu


def synth(n, d, kernel="se", seed=0):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    gp = GPOracle(1e-6, float(y.max() - y.min()), 0.25 * np.ones(d), float(y.mean()), kernel)
    gp.add_data(X, y)
    return gp


def slices_of(x, S):
    """Balanced base-256 digits (ozaki.cu header): X = rint(x * 127 * 2^48) as exact Python integers;
    byte k of X + sum_k 128 * 256^k, minus 128, is digit k; slice s is digit 6 - s."""
    X = np.rint(x * float(127 * 2 ** 48)).astype(np.int64)
    Y = X + 0x0080808080808080
    assert np.all(Y >= 0) and np.all(Y < 2 ** 56)
    return [((Y >> (8 * (6 - s))) & 255) - 128 for s in range(S)]


@pytest.mark.parametrize("n,d,S,mc,extra", [(64, 2, 2, 128, False), (256, 4, 3, 128, False), (300, 8, 5, 200, False),
                                            (512, 3, 7, 128, False), (200, 5, 6, 128, False), (130, 2, 7, 100, False),
                                            (256, 4, 4, 128, False), (300, 8, 5, 128, True), (192, 3, 3, 128, True),
                                            (128, 2, 7, 128, True)])
def test_slices_and_accumulators_exact(ctx, n, d, S, mc, extra):
    gp = synth(n, d, seed=n)
    ctx.fit("se", gp.X, gp.Y, gp.ell[None], [gp.rho], [gp.sn2], [gp.bias])
    Xc = qmc.Sobol(d=d, scramble=False).random(256)[:mc]
    out = ctx.ozaki_debug(S, Xc, extra=extra)
    npad = out["ws"].shape[1]
    W = np.zeros((npad, npad))
    W[:n, :n] = ctx.factor("W")
    W[np.arange(n, npad), np.arange(n, npad)] = 1.0
    mx = np.max(np.abs(W), axis=1)
    _, e = np.frexp(mx)
    assert np.allclose(out["rowscale"], gp.rho * np.exp2(e) / 127.0 ** 2, rtol=1e-15, atol=0)
    ref_ws = slices_of(W * np.exp2(-e)[:, None], S)
    for s in range(S):
        assert np.array_equal(out["ws"][s].astype(np.int64), ref_ws[s]), "W slice %d" % s
    # K* slices: the device exp may differ from NumPy's in the last ulp, so compare the value they encode
    kap = np.zeros((out["ks"].shape[1], npad))
    kap[:mc, :n] = kernel_matrix("se", Xc, gp.X, gp.ell, 1.0)
    enc = sum(out["ks"][s].astype(np.float64) * 256.0 ** -s for s in range(S)) / 127.0
    assert np.max(np.abs(enc - kap)) <= 0.51 * 256.0 ** -(S - 1) / 127.0 + 1e-15
    # TMEM accumulators of tile 0: exact integer contraction of the returned slices
    ks, ws = out["ks"].astype(np.int64), out["ws"].astype(np.int64)
    G = S - 1 + (1 if extra else 0)
    for rb in range(npad // 64):
        kmax = (rb + 1) * 64
        rows = slice(rb * 64, rb * 64 + 64)
        for g in range(G + 1):
            ref = sum(ks[g - s][:128, :kmax] @ ws[s][rows, :kmax].T for s in range(g + 1) if s < S and g - s < S)
            assert np.array_equal(out["acc"][rb, g].astype(np.int64), ref), (rb, g)
    # reassembly + reductions
    v = np.zeros((128, npad))
    for rb in range(npad // 64):
        a = np.zeros((128, 64))
        for g in range(G, -1, -1):
            a = a * 2.0 ** -8 + out["acc"][rb, g]
        v[:, rb * 64:rb * 64 + 64] = a * out["rowscale"][rb * 64:rb * 64 + 64]
    k = min(mc, 128)
    assert np.allclose(out["s2"][:k], gp.rho - np.sum(v[:k] ** 2, axis=1), rtol=1e-12, atol=1e-12)
    # the mean does not go through the int8 contraction: mu = bias + k*^T beta in FP64 inside the slicer
    mu64, _ = ctx.predict(Xc[:k])
    # (k*^T beta cancels like cond(K): a few 1e-9 on the tiny ill-conditioned fits used here)
    assert np.max(np.abs(out["mu"][:k] - mu64)) < 2e-8 * max(1.0, np.max(np.abs(mu64)))


@pytest.mark.parametrize("n,d,kernel", [(1024, 8, "se"), (700, 8, "matern52"), (2048, 8, "se")])
def test_ozaki_scores_match_oracle(ctx, n, d, kernel):
    gp = synth(n, d, kernel, seed=n + 1)
    ctx.fit(kernel, gp.X, gp.Y, gp.ell[None], [gp.rho], [gp.sn2], [gp.bias])
    Xc = qmc.Sobol(d=d, scramble=False).random_base2(12)[:3000]
    target = float(gp.predict(gp.X)[0].max())
    ref = gp.get_improvement(target, Xc)
    mu, s2 = gp.predict(Xc)
    f64val, _, f64best = ctx.score(1, target, Xc, want_best=True)
    ctx.set_precision(1, 1e-8)                      # BO_PREC_OZAKI, error-model driven slice count
    val, _, best = ctx.score(1, target, Xc, want_best=True)
    gmu, gs2 = ctx.predict(Xc)
    assert rel_err(gmu, mu) < 1e-6 and rel_err(gs2, s2, 1e-9) < 1e-6
    assert rel_err(val, ref, 1e-9) < 1e-6
    assert best[1] == int(np.argmax(ref)) == f64best[1]
    # gradients are still served (FP64 path) while the int8 path is selected
    v1, g1, _ = ctx.score(1, target, Xc[:5], grad=True)
    rv, rg = gp.get_improvement(target, Xc[:5], grad=True)
    assert rel_err(v1, rv, 1e-9) < 1e-6 and rel_err(g1, rg, 1e-9) < 1e-5
    ctx.set_precision(0, 1e-9)
    back, _, _ = ctx.score(1, target, Xc)
    assert np.array_equal(back, f64val)


def test_ozaki_slice_count_follows_tolerance(ctx):
    gp = synth(512, 8, seed=3)
    ctx.fit("se", gp.X, gp.Y, gp.ell[None], [gp.rho], [gp.sn2], [gp.bias])
    Xc = qmc.Sobol(d=8, scramble=False).random_base2(9)
    mu, s2 = gp.predict(Xc)
    errs = []
    ctx.set_rescue(False)                           # raw slice error: no FP64 rescue of the candidates it flags
    for S in (3, 4, 5, 6):
        ctx.set_precision(1, float(S))              # tol >= 2 pins the slice count
        gmu, gs2 = ctx.predict(Xc)
        errs.append(max(np.max(np.abs(gmu - mu)), np.max(np.abs(gs2 - s2))))
    assert errs[0] > errs[1] > errs[2] and errs[3] <= errs[2]
    assert errs[3] < 1e-10 and errs[1] / errs[2] > 60   # ~2^8 per extra slice until the FP64 floor


def test_full_size_headline_shape_int8_vs_fp64_and_oracle(ctx):
    """BASELINE headline shape (RBF n=4096, d=8, EI) at bench.py's default tolerance (1e-8 -> 5 base-256
    slices selected, tiers on): the int8-slice path against the FP64 path on 2^18 candidates (many
    32768-candidate chunks) and against the oracle on a slice; identical arg max and top-10."""
    gp = synth(4096, 8, "se", seed=0)
    ctx.fit("se", gp.X, gp.Y, gp.ell[None], [gp.rho], [gp.sn2], [gp.bias])
    Xc = qmc.Sobol(d=8, scramble=False).random_base2(18)
    target = float(gp.predict(gp.X)[0].max())
    f64val, _, f64best = ctx.score(1, target, Xc, want_best=True)
    top64 = ctx.topk(10)
    ctx.set_precision(1, 1e-8)
    val, _, best = ctx.score(1, target, Xc, want_best=True)
    top8 = ctx.topk(10)
    # 1e-8 selects 5 slices; a pass of >= 8 chunks pilots 4096 candidates one half-level down (4 slices + extra) and
    # keeps that level when at most 10 % of them are flagged (test_gpu_rescue.py::test_tiered_levels... pins each case)
    t = ctx.tier_info()
    assert ctx.precision_info()[0] == 1 and t["first"] == (4, True) and t["rest"] in ((4, True), (5, False)), t
    assert rel_err(val, f64val) < 5e-7                  # the rescue tolerance; measured 3e-8 .. 6e-8; floor 1e-12 max
    assert best[1] == f64best[1] and np.array_equal(top8[0], top64[0])
    sl = slice(32700, 32900)
    ref = gp.get_improvement(target, Xc[sl])
    assert rel_err(val[sl], ref) < 1e-6
    # the mean never goes through the int8 contraction
    mu, s2 = ctx.predict(Xc[:5000])
    rmu, rs2 = gp.predict(Xc[:5000])
    assert rel_err(mu, rmu) < 1e-9 and rel_err(s2, rs2, 1e-9) < 1e-7
    # size-independent properties: 0 < s2 <= rho, scores independent of chunking
    assert np.all(s2 > 0) and np.all(s2 <= gp.rho * (1 + 1e-9))
    again, _, _ = ctx.score(1, target, Xc[sl])
    assert rel_err(again, val[sl]) < 5e-7               # tiers on: the level, and with it the last digits, depends on the pass length
    ctx.set_option("oz_tiered", 0)                      # one level per pass: bit-identical whatever the chunking
    full, _, _ = ctx.score(1, target, Xc)
    again, _, _ = ctx.score(1, target, Xc[sl])
    ctx.set_option("oz_tiered", 1)
    assert ctx.tier_info()["rest"] == (5, False)
    assert np.array_equal(again, full[sl])
    # the cheaper level (4 slices + first dropped pair group, 13 digit pairs) still meets 1e-6
    ctx.set_precision(1, 4.5)
    fast, _, fbest = ctx.score(1, target, Xc, want_best=True)
    assert ctx.precision_info() == (1, 4, True) and fbest[1] == f64best[1]
    assert rel_err(fast, f64val) < 5e-7                 # 4.2e-7 raw at this level; the flagged candidates are FP64 now
    assert ctx.rescue_info()[1] > 0


@pytest.mark.parametrize("d,m,ndraw,M", [(2, 128, 7, 1500), (16, 500, 300, 4097), (8, 1024, 256, 40000), (5, 33, 64, 1024)])
def test_thompson_batch_int8_path(d, m, ndraw, M):
    """Thompson draws on a shared basis (BASELINE config 4 family) through the int8-slice contraction
    (sliced cosine features x sliced Theta on tcgen05): values and per-draw first arg max against the
    NumPy restatement and the FP64 tensor-core path."""
    from pybo_b200 import models
    gp = synth(80, d, "se", seed=3)
    mine = models.GP(1e-3, gp.rho, gp.ell, gp.bias)
    mine.add_data(gp.X, gp.Y)
    tb = models.ThompsonBatch(mine, m=m, ndraw=ndraw, rng=5)
    Xc = qmc.Sobol(d=d, scramble=False).random_base2(int(np.ceil(np.log2(M))))[:M]
    ref = (tb.bias + (tb.scale * np.cos(Xc @ tb.W[0].T + tb.b[0])) @ tb.theta.T).T
    F64 = tb.get(Xc)
    bv64, bi64 = tb.argmax(Xc)
    tb.set_precision("int8", 1e-8)
    F8 = tb.get(Xc)
    bv8, bi8 = tb.argmax(Xc)
    assert tb._context().launch_count() > 0
    # draws change sign, so the error is measured against the draws' scale (max |F|), not element by element:
    # tol = 1e-8 is the int8 path's target relative to the prior standard deviation sqrt(rho)
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(F8 - ref)) < 1e-7 * scale and np.max(np.abs(F8 - F64)) < 1e-7 * scale
    assert np.array_equal(bi8, np.argmax(ref, axis=1)) and np.array_equal(bi8, bi64)
    assert np.allclose(bv8, F8[np.arange(ndraw), bi8], rtol=0, atol=0)
    # pinned levels: error shrinks by ~2^8 per slice
    errs = []
    for lvl in (3.0, 4.0, 5.0):
        tb.set_precision("int8", lvl)
        errs.append(np.max(np.abs(tb.get(Xc) - ref)))
    assert errs[0] > errs[1] > errs[2] and errs[0] / errs[1] > 50


@pytest.mark.parametrize("cl", [2, 4])
def test_cluster_multicast_variant_is_bit_identical(ctx, cl):
    """The thread-block-cluster form of the contraction (K* tile fetched once per cluster by TMA multicast, stage
    release committed to every CTA of the cluster) computes exactly what the single-CTA form computes."""
    gp = synth(1024, 8, "se", seed=1)
    ctx.fit("se", gp.X, gp.Y, gp.ell[None], [gp.rho], [gp.sn2], [gp.bias])
    Xc = qmc.Sobol(d=8, scramble=False).random_base2(15)[:33000]
    target = float(gp.predict(gp.X)[0].max())
    ctx.set_rescue(False)
    for level in (5.0, 4.5):
        ctx.set_precision(1, level)
        ctx.set_option("oz_cluster", 1)
        base, _, b0 = ctx.score(1, target, Xc, want_best=True)
        ctx.set_option("oz_cluster", cl)
        got, _, b1 = ctx.score(1, target, Xc, want_best=True)
        assert np.array_equal(got, base) and b0 == b1
    ctx.set_option("oz_cluster", 0)
