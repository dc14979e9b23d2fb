"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the
float64 oracle on identical seeded inputs and against the committed golden
vectors.

Tolerance (BASELINE.json north_star): outputs within 1e-6 relative of the
reference path and identical arg-max index.  The metric is
max |got - ref| / max(|ref|, floor) with floor = 1e-12 * max|ref| for mu and the
acquisition values (conftest.rel_err).  For the predictive variance the floor is
1e-9 * rho: s2 = rho - |v|^2 cancels to O(sn2) at the data, where two float64
LAPACK builds already disagree at that level.
"""

import glob
import os
import pickle

import numpy as np
import pytest
from scipy.stats import qmc

from conftest import rel_err
from oracle import GPOracle, MixtureOracle, FourierSampleOracle, ucb_beta, ucb_index, kernel_matrix

pytestmark = pytest.mark.gpu

TOL = 1e-6


def synth(n, d, kernel, seed=0, sn2=1e-6):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    gp = GPOracle(sn2, float(y.max() - y.min()), 0.25 * np.ones(d), float(y.mean()), kernel)
    gp.add_data(X, y)
    return gp


def fit_ctx(ctx, gp):
    ctx.fit(gp.kernel, gp.X, gp.Y, gp.ell[None], [gp.rho], [gp.sn2], [gp.bias])
    return ctx


def sobol(M, d):
    return qmc.Sobol(d=d, scramble=False).random_base2(int(np.ceil(np.log2(M))))[:M]


# ---- Gram / Cholesky / factors ------------------------------------------------
@pytest.mark.parametrize("kernel", ["se", "matern52"])
@pytest.mark.parametrize("n,d", [(1, 1), (37, 3), (128, 2), (300, 8), (513, 16)])
def test_gram_matches_oracle(ctx, kernel, n, d):
    rng = np.random.RandomState(n)
    X, ell = rng.rand(n, d), 0.2 + 0.3 * rng.rand(d)
    K = ctx.gram(kernel, X, ell, 1.7, 1e-3)
    ref = kernel_matrix(kernel, X, X, ell, 1.7) + 1e-3 * np.eye(n)
    assert rel_err(K, ref) < 1e-12


@pytest.mark.parametrize("n", [1, 5, 64, 65, 200, 512, 1000])
def test_cholesky_matches_lapack(ctx, n):
    rng = np.random.RandomState(n)
    B = rng.randn(n, n + 3)
    A = B @ B.T + 0.5 * np.eye(n)
    L = ctx.cholesky(A)
    ref = np.linalg.cholesky(A)
    assert np.all(np.triu(L, 1) == 0)
    assert np.max(np.abs(L - ref)) < 1e-11 * np.max(np.abs(ref))
    assert np.max(np.abs(L @ L.T - A)) < 1e-12 * np.max(np.abs(A)) * n


def test_cholesky_batched_and_not_pd(ctx):
    rng = np.random.RandomState(0)
    A = np.array([(lambda B: B @ B.T + np.eye(96))(rng.randn(96, 100)) for _ in range(5)])
    L = ctx.cholesky(A)
    for b in range(5):
        assert np.max(np.abs(L[b] - np.linalg.cholesky(A[b]))) < 1e-10
    bad = np.eye(70)
    bad[40, 40] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        ctx.cholesky(bad)


@pytest.mark.parametrize("batch,n", [(10, 2048), (24, 1024), (5, 1500)])
def test_cholesky_many_matrices_repeated_calls(ctx, batch, n):
    """Many matrices at once: sub-batches on their own stream pairs, diagonal / panel as separate launches, trailing
    updates over groups of four panels; the same device buffers three times in a row, so that the third call replays
    the captured CUDA graph where the shape allows one (10 x 2048: graph + two stream pairs inside the capture)."""
    import torch
    rng = np.random.RandomState(batch)
    B = rng.randn(n, n + 8)
    A0 = B @ B.T / n + 0.5 * np.eye(n)
    A = np.stack([A0 + 0.01 * b * np.eye(n) for b in range(batch)])
    ref = [np.linalg.cholesky(A[b]) for b in (0, batch // 2, batch - 1)]
    src = torch.from_numpy(A).cuda()
    work = torch.empty_like(src)
    for rep in range(3):
        work.copy_(src)
        torch.cuda.synchronize()
        info = ctx.cholesky_device(n, batch, work.data_ptr())
        assert not info.any()
        for r, b in zip(ref, (0, batch // 2, batch - 1)):
            L = np.tril(work[b].cpu().numpy())
            assert np.max(np.abs(L - r)) < 1e-11 * np.max(np.abs(r)), (rep, b)


@pytest.mark.parametrize("kernel,n,d", [("se", 40, 2), ("matern52", 333, 5), ("se", 1024, 4)])
def test_fit_factors(ctx, kernel, n, d):
    gp = synth(n, d, kernel, seed=n)
    fit_ctx(ctx, gp)
    L = ctx.factor("L")
    assert np.max(np.abs(L - gp.L)) < 1e-9 * np.max(np.abs(gp.L))
    W = ctx.factor("W")
    assert np.max(np.abs(W @ gp.L - np.eye(n))) < 1e-7
    assert rel_err(ctx.factor("alpha"), gp.alpha, 1e-9) < 1e-6
    assert rel_err(ctx.factor("beta"), gp.beta, 1e-9) < 1e-5
    ll = ctx.loglik()[0]
    assert abs(ll - gp.loglikelihood()) < 1e-8 * abs(gp.loglikelihood())


# ---- golden vectors --------------------------------------------------------------
def test_golden_posterior_vectors(ctx, golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "posterior_*.npz")))
    assert files
    for f in files:
        g = np.load(f)
        rho = float(g["rho"])
        ctx.fit(str(g["kernel"]), g["X"], g["Y"], g["ell"][None], [rho], [float(g["sn2"])], [float(g["bias"])])
        mu, s2, dmu, ds2 = ctx.predict(g["Xc"], grad=True)
        assert rel_err(mu, g["mu"]) < TOL, f
        assert np.max(np.abs(s2 - g["s2"]) / np.maximum(np.abs(g["s2"]), 1e-9 * rho)) < TOL, f
        assert rel_err(dmu, g["dmu"], 1e-9) < TOL and rel_err(ds2, g["ds2"], 1e-9) < TOL, f
        for acq, param, key, gkey in ((1, float(g["target"]), "ei", "dei"),
                                      (2, float(g["target"]) + 0.05, "pi", "dpi"),
                                      (3, float(g["beta"]), "ucb", "ducb")):
            val, grad, best = ctx.score(acq, param, g["Xc"], grad=True, want_best=True)
            assert rel_err(val, g[key], 1e-9) < TOL, (f, key)
            assert rel_err(grad, g[gkey], 1e-9) < 10 * TOL, (f, key)
            assert best[1] == int(np.argmax(g[key])) and best[0] == val[best[1]], (f, key)


def test_golden_mixture_vectors(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "mixture_se_n150_d3_s5.npz"))
    ctx.fit("se", g["X"], g["Y"], g["ell"], g["rho"], g["sn2"], g["bias"])
    mu, s2, dmu, ds2 = ctx.predict(g["Xc"], grad=True)
    assert rel_err(mu, g["mu"]) < TOL and rel_err(s2, g["s2"], 1e-9) < TOL
    assert rel_err(dmu, g["dmu"], 1e-9) < TOL and rel_err(ds2, g["ds2"], 1e-9) < TOL
    ei, dei, best = ctx.score(1, float(g["target"]), g["Xc"], grad=True, want_best=True)
    assert rel_err(ei, g["ei"], 1e-9) < TOL and rel_err(dei, g["dei"], 1e-9) < 10 * TOL
    assert best[1] == int(np.argmax(g["ei"]))


# ---- scoring against the oracle on seeded inputs -----------------------------------
@pytest.mark.parametrize("kernel,n,d,M", [("se", 1024, 4, 5000), ("matern52", 700, 8, 3000),
                                          ("se", 20, 2, 1), ("se", 129, 16, 257), ("matern52", 64, 1, 130)])
def test_scores_match_oracle(ctx, kernel, n, d, M):
    gp = synth(n, d, kernel, seed=n + d)
    fit_ctx(ctx, gp)
    Xc = sobol(M, d)
    mu, s2 = gp.predict(Xc)
    target = float(gp.predict(gp.X)[0].max())
    gmu, gs2 = ctx.predict(Xc)
    assert rel_err(gmu, mu) < TOL
    assert np.max(np.abs(gs2 - s2) / np.maximum(np.abs(s2), 1e-9 * gp.rho)) < TOL
    for acq, ref in ((0, mu), (1, gp.get_improvement(target, Xc)), (2, gp.get_tail(target + 0.05, Xc)),
                     (3, ucb_index(ucb_beta(n), mu, s2))):
        val, _, best = ctx.score(acq, {0: 0.0, 1: target, 2: target + 0.05, 3: ucb_beta(n)}[acq], Xc, want_best=True)
        assert rel_err(val, ref, 1e-9) < TOL, acq
        assert best[1] == int(np.argmax(ref)), acq


@pytest.mark.parametrize("kernel,n,d,M", [("se", 2100, 32, 700), ("matern52", 150, 31, 300), ("se", 8190, 8, 600)])
def test_edge_shapes_max_dimension_and_large_ragged_n(ctx, kernel, n, d, M):
    """Largest supported input dimension (32, and 31 padded to it) and a large factor whose order is not a
    multiple of the 128 padding (n = 8190): both precision paths against the oracle."""
    rng = np.random.RandomState(n + d)
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    # lengthscale grows with sqrt(d) so that candidates stay correlated with the data in 32 dimensions
    gp = GPOracle(1e-4, float(y.max() - y.min()), 0.25 * max(1.0, np.sqrt(d / 8.0)) * np.ones(d), float(y.mean()), kernel)
    gp.add_data(X, y)
    fit_ctx(ctx, gp)
    Xc = sobol(M, d)
    mu, s2 = gp.predict(Xc)
    target = float(gp.predict(gp.X[:200])[0].max())
    ref = gp.get_improvement(target, Xc)
    gmu, gs2 = ctx.predict(Xc)
    assert rel_err(gmu, mu) < TOL and np.max(np.abs(gs2 - s2) / np.maximum(np.abs(s2), 1e-9 * gp.rho)) < TOL
    val, _, best = ctx.score(1, target, Xc, want_best=True)
    assert rel_err(val, ref, 1e-9) < TOL and best[1] == int(np.argmax(ref))
    ctx.set_precision(1, 1e-9)
    v8, _, b8 = ctx.score(1, target, Xc, want_best=True)
    assert rel_err(v8, ref, 1e-8) < TOL and b8[1] == best[1]
    v1, g1, _ = ctx.score(1, target, Xc[:3], grad=True)
    rv, rg = gp.get_improvement(target, Xc[:3], grad=True)
    assert rel_err(v1, rv, 1e-9) < TOL and rel_err(g1, rg, 1e-9) < 10 * TOL


def test_device_sobol_grid_bit_exact_and_staged_scoring(ctx):
    """bo_candidates_sobol: the device generator reproduces SciPy's unscrambled Sobol points bit for bit
    (any offset), and scoring the grid staged in the handle equals scoring the same points from the host."""
    from scipy.stats import qmc
    from pybo_b200 import _lib, models, policies, solvers
    for d in (1, 4, 16, 32):
        ref = qmc.Sobol(d=d, scramble=False).random_base2(13)
        assert np.array_equal(ctx.sobol(d, 0, 8192), ref)
        assert np.array_equal(ctx.sobol(d, 1000, 3001), ref[1000:4001])
    b = np.array([[-5, 10.0], [0, 15]])
    assert np.allclose(ctx.sobol(2, 5, 100, b), _lib.sobol_points(2, np.arange(5, 105), b), rtol=1e-15, atol=0)
    with pytest.raises(ValueError):
        ctx.sobol(3, 2 ** 30 - 10, 100)
    gp = synth(400, 4, "se", seed=12)
    fit_ctx(ctx, gp)
    target = float(gp.predict(gp.X)[0].max())
    pts = ctx.sobol(4, 64, 20000)
    val, _, best = ctx.score(1, target, pts, want_best=True)
    ctx.sobol(4, 64, 20000, out="staged")
    sval, sbest = ctx.score_staged(1, target, 20000, want_values=True)
    assert np.array_equal(sval, val) and sbest == best
    with pytest.raises(_lib.BackendError):
        ctx.score_staged(1, target, 19999)                 # the handle holds a different grid
    ctx.score(1, target, pts[:10])                          # a host-pointer call un-stages the grid
    with pytest.raises(_lib.BackendError):
        ctx.score_staged(1, target, 20000)
    # through the plugin surface: solver(grid='sobol') on a device model = the same on the oracle model
    gpu = models.make_gp(gp.sn2, gp.rho, gp.ell, gp.bias)
    gpu.add_data(gp.X, gp.Y)
    bounds = np.array([[0, 1.0]] * 4)
    a, r = policies.EI(gpu, bounds, list(gp.X)), policies.EI(gp, bounds, list(gp.X))
    p3, v3, i3 = a.best_of_sobol(bounds, 4096, 3)
    rv = r(_lib.sobol_points(4, np.arange(4096), bounds))
    assert np.array_equal(i3, np.lexsort((np.arange(4096), -rv))[:3]) and rel_err(v3, rv[i3], 1e-9) < TOL
    xa, fa = solvers.solve_lbfgs(a, bounds, ngrid=4096, grid="sobol")
    xb, fb = solvers.solve_lbfgs(r, bounds, ngrid=4096, grid="sobol")
    assert np.allclose(xa, xb, atol=1e-4) and abs(fa - fb) <= 1e-6 * max(1, abs(fb))


def test_predict_at_training_points_and_incumbent_target(ctx):
    """policies/simple.py:21: target = max_i mu(x_i); recommenders.py:34-35."""
    gp = synth(500, 4, "se", seed=9)
    fit_ctx(ctx, gp)
    mu, s2 = gp.predict(gp.X)
    gmu, gs2 = ctx.predict(gp.X)
    assert rel_err(gmu, mu) < TOL
    assert np.max(np.abs(gs2 - s2)) < 1e-9 * gp.rho
    assert int(np.argmax(gmu)) == int(np.argmax(mu))


def test_ties_resolve_to_first_index(ctx):
    """EI underflows to exactly 0 far from the data; arg max must be the first index."""
    gp = GPOracle(1e-6, 1.0, [0.01, 0.01], 0.0, "se")
    gp.add_data(np.array([[0.5, 0.5], [0.52, 0.5]]), np.array([1.0, 0.9]))
    fit_ctx(ctx, gp)
    Xc = np.random.RandomState(0).rand(4000, 2) * 0.2            # all far away: EI == 0
    val, _, best = ctx.score(1, 50.0, Xc, want_best=True)          # z = -50: phi, Phi underflow to 0
    assert np.all(val == 0.0) and best == (0.0, 0)
    Xc[1234] = Xc[77]
    val, _, best = ctx.score(0, 0.0, Xc, want_best=True)
    assert best[1] == int(np.argmax(val))
    idx, tv = ctx.topk(10)
    order = np.lexsort((np.arange(len(val)), -val))[:10]
    assert np.array_equal(idx, order) and np.array_equal(tv, val[order])


def test_topk_matches_argsort(ctx):
    gp = synth(300, 3, "se", seed=4)
    fit_ctx(ctx, gp)
    Xc = sobol(20000, 3)
    val, _, _ = ctx.score(3, ucb_beta(300), Xc)
    idx, tv = ctx.topk(10)
    order = np.argsort(val)[::-1][:10]                              # lbfgs.py:51
    assert np.array_equal(idx, order) and np.array_equal(tv, val[order])


def test_mixture_matches_oracle(ctx):
    rng = np.random.RandomState(5)
    base = synth(260, 4, "matern52", seed=5, sn2=1e-4)
    gps = []
    for _ in range(6):
        g = GPOracle(base.sn2 * np.exp(0.3 * rng.randn()), base.rho * np.exp(0.2 * rng.randn()),
                     base.ell * np.exp(0.2 * rng.randn(4)), base.bias + 0.05 * rng.randn(), "matern52")
        g.add_data(base.X, base.Y)
        gps.append(g)
    mix = MixtureOracle(gps)
    ctx.fit("matern52", base.X, base.Y, [g.ell for g in gps], [g.rho for g in gps], [g.sn2 for g in gps],
            [g.bias for g in gps])
    Xc = sobol(1500, 4)
    mu, s2 = mix.predict(Xc)
    target = float(mix.predict(base.X)[0].max())
    gmu, gs2 = ctx.predict(Xc)
    assert rel_err(gmu, mu) < TOL and rel_err(gs2, s2, 1e-9) < TOL
    for acq, param, ref in ((1, target, mix.get_improvement(target, Xc)), (2, target, mix.get_tail(target, Xc)),
                            (3, 4.0, ucb_index(4.0, mu, s2))):
        val, _, best = ctx.score(acq, param, Xc, want_best=True)
        assert rel_err(val, ref, 1e-9) < TOL and best[1] == int(np.argmax(ref))
    assert np.allclose(ctx.loglik(), [g.loglikelihood() for g in gps], rtol=1e-8)


def test_single_point_gradient_calls(ctx):
    """The L-BFGS callback shape: f(x[None], grad=True) (solvers/lbfgs.py:56-58)."""
    gp = synth(150, 3, "se", seed=6, sn2=1e-4)
    fit_ctx(ctx, gp)
    target = float(gp.predict(gp.X)[0].max())
    for x in np.random.RandomState(1).rand(5, 3):
        val, grad, _ = ctx.score(1, target, x[None], grad=True)
        ref, rgrad = gp.get_improvement(target, x[None], grad=True)
        assert rel_err(val, ref, 1e-9) < TOL and rel_err(grad, rgrad, 1e-9) < 10 * TOL


@pytest.mark.parametrize("kernel,n,d", [("se", 300, 4), ("matern52", 1030, 8)])
def test_small_batch_gradient_calls(ctx, kernel, n, d):
    """Batches of 2..17 points with gradients (the batched multi-start refinement): every small-batch
    code path (1/2/4 right-hand sides per warp, 8/16 staged in shared memory, tiled GEMM above 16)."""
    gp = synth(n, d, kernel, seed=n, sn2=1e-5)
    fit_ctx(ctx, gp)
    target = float(gp.predict(gp.X)[0].max())
    rng = np.random.RandomState(2)
    for M in (2, 3, 4, 5, 8, 10, 13, 16, 17):
        X = rng.rand(M, d)
        val, grad, _ = ctx.score(1, target, X, grad=True)
        ref, rgrad = gp.get_improvement(target, X, grad=True)
        assert rel_err(val, ref, 1e-9) < TOL and rel_err(grad, rgrad, 1e-9) < 10 * TOL, M
        mu, s2, dmu, ds2 = ctx.predict(X, grad=True)
        rmu, rs2, rdmu, rds2 = gp.predict(X, grad=True)
        assert rel_err(mu, rmu) < TOL and rel_err(s2, rs2, 1e-9) < TOL
        assert rel_err(dmu, rdmu, 1e-9) < 10 * TOL and rel_err(ds2, rds2, 1e-9) < 10 * TOL


# ---- Thompson ---------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["se", "matern52"])
def test_thompson_draw_matches_oracle(ctx, kernel):
    from pybo_b200 import models
    gp = synth(60, 3, kernel, seed=8, sn2=1e-3)
    ref = FourierSampleOracle(gp, 200, rng=21)
    mine = models.GP(gp.sn2, gp.rho, gp.ell, gp.bias, kernel=kernel)
    mine.add_data(gp.X, gp.Y)
    draw = mine.sample_f(200, rng=21)
    assert np.allclose(draw.theta, ref.theta, rtol=1e-8, atol=1e-11)     # theta is built on the device (bo_thompson_build)
    Xc = sobol(3000, 3)
    F, G = draw.get(Xc, grad=True)
    RF, RG = ref.get(Xc, grad=True)
    assert rel_err(F, RF, 1e-9) < TOL and rel_err(G, RG, 1e-9) < TOL
    bv, bi = draw.argmax(Xc)
    assert bi == int(np.argmax(RF)) and bv == F[bi]


@pytest.mark.parametrize("d,m,ndraw,M", [(2, 128, 7, 1000), (16, 500, 300, 4097), (5, 33, 256, 64)])
def test_thompson_batch_shared_basis(ctx, d, m, ndraw, M):
    """BASELINE config 4 shape family: ndraw posterior draws x M candidates on one random-feature
    basis (tensor-core contraction with on-the-fly cosine features), per-draw arg max."""
    from pybo_b200 import models
    gp = synth(80, d, "se", seed=3, sn2=1e-3)
    mine = models.GP(gp.sn2, gp.rho, gp.ell, gp.bias)
    mine.add_data(gp.X, gp.Y)
    tb = models.ThompsonBatch(mine, m=m, ndraw=ndraw, rng=5)
    Xc = sobol(M, d)
    F = tb.get(Xc)
    ref = tb.bias + (tb.scale * np.cos(Xc @ tb.W[0].T + tb.b[0])) @ tb.theta.T
    assert rel_err(F, ref.T, 1e-9) < TOL
    bv, bi = tb.argmax(Xc)
    assert np.array_equal(bi, np.argmax(ref, axis=0))


# ---- the plugin surface on the GPU model ---------------------------------------------------
def _branin(x):
    x = np.array(x, ndmin=2)
    y = (x[:, 1] - (5.1 / (4 * np.pi ** 2)) * x[:, 0] ** 2 + 5 * x[:, 0] / np.pi - 6) ** 2
    y += 10 * (1 - 1 / (8 * np.pi)) * np.cos(x[:, 0]) + 10
    return float(-np.squeeze(y / 10.0))


def test_policies_solver_recommender_on_gpu_model():
    from pybo_b200 import models, policies, recommenders, solvers
    rng = np.random.RandomState(0)
    bounds = np.array([[-5, 10.0], [0, 15]])
    X = bounds[:, 0] + (bounds[:, 1] - bounds[:, 0]) * rng.rand(12, 2)
    Y = np.array([_branin(x) for x in X])
    ell = 0.25 * (bounds[:, 1] - bounds[:, 0])
    ref = GPOracle(1e-6, 10.0, ell, -5.0, "se")
    ref.add_data(X, Y)
    gpu = models.make_gp(1e-6, 10.0, ell, -5.0)
    gpu.add_data(list(X), list(Y))
    grid = bounds[:, 0] + (bounds[:, 1] - bounds[:, 0]) * rng.rand(4000, 2)
    for name in ("EI", "PI", "UCB"):
        a, b = getattr(policies, name)(gpu, bounds, list(X)), getattr(policies, name)(ref, bounds, list(X))
        assert abs(a.param - b.param) <= 1e-6 * max(1.0, abs(b.param))
        va, ga = a(grid, grad=True)
        vb, gb = b(grid, grad=True)
        assert rel_err(va, vb, 1e-9) < TOL and rel_err(ga, gb, 1e-9) < 10 * TOL
        assert int(np.argmax(va)) == int(np.argmax(vb))
        idx, val = a.best_of(grid, 10)
        assert np.array_equal(idx, np.lexsort((np.arange(len(vb)), -vb))[:10])
        xa, fa = solvers.solve_lbfgs(a, bounds, xgrid=grid)
        xb, fb = solvers.solve_lbfgs(b, bounds, xgrid=grid)
        assert np.allclose(xa, xb, atol=1e-4 * np.max(bounds[:, 1] - bounds[:, 0])) and abs(fa - fb) <= 1e-6 * max(1, abs(fb))
        # batched multi-start refinement on the device model vs the sequential SciPy runs on the oracle
        xc, fc = solvers.solve_lbfgs_batched(a, bounds, xgrid=grid)
        xd, fd = solvers.solve_lbfgs(b, bounds, xgrid=grid, pick="best")
        assert abs(fc - fd) <= 1e-6 * max(1, abs(fd)) and abs(float(b(xc[None])[0]) - fc) <= 1e-6 * max(1, abs(fc))
    assert np.allclose(recommenders.best_latent(gpu, bounds, list(X)), recommenders.best_latent(ref, bounds, list(X)), atol=1e-4)
    assert np.array_equal(recommenders.best_incumbent(gpu, bounds, list(X)), recommenders.best_incumbent(ref, bounds, list(X)))
    # copy() shares the fitted state; add_data on the copy leaves the original alone
    c = gpu.copy()
    assert c._fit is gpu._fit
    c.add_data(X[0] + 0.1, 0.0)
    assert gpu.ndata == 12 and c.ndata == 13 and gpu._fit is not None
    mu0 = gpu.predict(grid[:5])[0]
    g2 = pickle.loads(pickle.dumps(gpu))
    assert np.array_equal(g2.predict(grid[:5])[0], mu0)


def test_bayesopt_branin_config1_gpu_vs_golden(golden_dir):
    """BASELINE config 1 with the GPU model behind the unchanged plugin surface."""
    import pybo_b200
    from pybo_b200 import models
    g = np.load(os.path.join(golden_dir, "bayesopt_branin_ei_20.npz"))
    bounds = g["bounds"]
    model = models.make_gp(1e-6, 10.0, 0.25 * (bounds[:, 1] - bounds[:, 0]), -5.0)
    xbest, model, info = pybo_b200.solve_bayesopt(_branin, bounds, model=model, niter=19, policy="ei",
                                                  solver="lbfgs", recommender="latent", rng=0)
    assert info.x.shape == (20, 2)
    assert np.allclose(info.x[:4], g["x"][:4], atol=1e-4) and np.allclose(info.y[:4], g["y"][:4], atol=1e-4)
    assert info.y.max() > -0.1


def test_bayesopt_branin_config1_full_trace_on_gpu_model(golden_dir):
    """All 20 observations of BASELINE config 1 on the GPU model.  A free-running BO loop amplifies 1e-12
    differences (each query becomes data for the next fit), so the full trace is checked by teacher forcing: at
    every iteration the GPU model holds the GOLDEN history, the same policy / solver / recommender objects consume
    the same random stream as the golden run, and the proposal and the recommendation must match the golden ones."""
    from pybo_b200 import inits, models, policies, recommenders, solvers
    from pybo_b200.bayesopt import get_component
    from pybo_b200.utils import rstate
    g = np.load(os.path.join(golden_dir, "bayesopt_branin_ei_20.npz"))
    bounds, gx, gy, gbest = g["bounds"], g["x"], g["y"], g["xbest"]
    width = bounds[:, 1] - bounds[:, 0]
    rng = rstate(0)
    policy = get_component("ei", policies, rng)
    solver = get_component("lbfgs", solvers, rng, lstrip="solve_")
    recommender = get_component("latent", recommenders, rng, lstrip="best_")
    model = models.make_gp(1e-6, 10.0, 0.25 * width, -5.0)
    assert np.array_equal(inits.init_middle(bounds)[0], gx[0]) and abs(_branin(gx[0]) - gy[0]) < 1e-12
    model.add_data(gx[0], gy[0])
    X = [gx[0]]
    worst_x = worst_b = 0.0
    for i in range(len(gx) - 1):
        index = policy(model, bounds, X)
        x, fmax = solver(index, bounds)
        del index
        worst_x = max(worst_x, float(np.max(np.abs(x - gx[i + 1]) / width)))
        assert np.allclose(x, gx[i + 1], atol=2e-3 * width), (i, x, gx[i + 1])
        assert abs(_branin(x) - gy[i + 1]) < 1e-2 * max(1.0, abs(gy[i + 1])), i
        model.add_data(gx[i + 1], gy[i + 1])                     # golden history from here on
        xbest = recommender(model, bounds, X)
        worst_b = max(worst_b, float(np.max(np.abs(xbest - gbest[i]) / width)))
        assert np.allclose(xbest, gbest[i], atol=2e-3 * width), (i, xbest, gbest[i])
        X.append(gx[i + 1])
    assert model.ndata == 20
    print("config 1, 20 observations, teacher forced: worst |dx| / width %.2e (proposals) %.2e (recommendations)" % (worst_x, worst_b))


def test_bayesopt_loop_with_device_grid_batched_refinement_and_appends():
    """The BO loop (reference bayesopt.py:262-276) with the opt-in fast pieces together: Sobol grid generated and
    scored on the device, all starts refined in lockstep, `add_data` appending to the device factors in place --
    against the same loop driven by the oracle model (same grid, same solver), which refits from scratch."""
    import pybo_b200
    from pybo_b200 import models
    bounds = np.array([[-5, 10.0], [0, 15]])
    ell = 0.25 * (bounds[:, 1] - bounds[:, 0])
    rng = np.random.RandomState(3)
    X0 = bounds[:, 0] + (bounds[:, 1] - bounds[:, 0]) * rng.rand(6, 2)
    Y0 = np.array([_branin(x) for x in X0])
    solver = ("lbfgs_batched", {"grid": "sobol", "ngrid": 2048})
    out = []
    for make in (lambda: models.make_gp(1e-6, 10.0, ell, -5.0), lambda: GPOracle(1e-6, 10.0, ell, -5.0, "se")):
        m = make()
        m.add_data(list(X0), list(Y0))
        xbest, model, info = pybo_b200.solve_bayesopt(_branin, bounds, model=m, niter=8, policy="ei", solver=solver,
                                                      recommender="incumbent", rng=0)
        out.append((xbest, model, info))
    (xa, ma, ia), (xb, mb, ib) = out
    assert ma.ndata == mb.ndata == 6 + 9
    assert ma._fit is not None and ma._fit.ctx.n == 15        # factors kept on the device through the appends
    assert np.allclose(ia.x[:5], ib.x[:5], atol=1e-3) and np.allclose(ia.y[:5], ib.y[:5], atol=1e-3)
    assert ia.y.max() >= Y0.max() - 1e-9 and np.all(np.isfinite(ia.y))


def test_default_model_mcmc_runs():
    import pybo_b200
    bounds = np.array([[-5, 10.0], [0, 15]])
    xbest, model, info = pybo_b200.solve_bayesopt(_branin, bounds, niter=3, rng=1)
    assert len(model) == 10 and info.x.shape == (4, 2) and np.all(np.isfinite(info.y))
    assert model.ndata == 6 + 4


# ---- size-independent properties at BASELINE sizes --------------------------------------------
def test_full_size_properties_n4096_d8(ctx):
    """Headline shape (RBF n=4096 d=8): W L = I, interpolation at the data, s2 in [0, rho],
    chunk-boundary invariance of the scores, and agreement with the oracle on a slice."""
    gp = synth(4096, 8, "se", seed=0)
    fit_ctx(ctx, gp)
    W, L = ctx.factor("W"), ctx.factor("L")
    E = W[:512] @ L[:, :512]
    assert np.max(np.abs(E[:, :512] - np.eye(512))) < 1e-8
    mu_d, s2_d = ctx.predict(gp.X[:2048])
    assert np.max(np.abs(mu_d - gp.Y[:2048])) < 1e-3 and np.all(s2_d < 1e-4 * gp.rho) and np.all(s2_d > -1e-9)
    Xc = sobol(20000, 8)
    target = float(gp.predict(gp.X)[0].max())
    val, _, best = ctx.score(1, target, Xc, want_best=True)
    gmu, gs2 = ctx.predict(Xc)
    assert np.all(gs2 > 0) and np.all(gs2 <= gp.rho * (1 + 1e-12))
    sl = slice(8000, 8400)                                  # straddles the 8192-candidate chunk edge
    ref = gp.get_improvement(target, Xc[sl])
    assert rel_err(val[sl], ref, 1e-9) < TOL
    val2, _, _ = ctx.score(1, target, Xc[sl])
    assert np.array_equal(val2, val[sl])                    # a candidate's score does not depend on its chunk
    assert best[1] == int(np.argmax(val))


# ---- error behaviour at the boundary -------------------------------------------------------
def test_error_codes_map_to_python_exceptions(ctx):
    """C ABI status codes -> the exceptions the reference path would raise: a failed factorisation is a
    numpy LinAlgError (what scipy.linalg.cholesky raises inside reggie), bad arguments are ValueError,
    calls before the state they need exists are BackendError."""
    from pybo_b200 import _lib
    rng = np.random.RandomState(0)
    X = rng.rand(10, 2)
    with pytest.raises(_lib.BackendError):
        ctx.n, ctx.d, ctx.S = 10, 2, 1
        ctx.predict(X)                                             # before any fit
    with pytest.raises(ValueError):
        ctx.fit("se", X, rng.rand(10), [[0.3, -0.1]], [1.0], [1e-6], [0.0])    # negative length scale
    with pytest.raises(ValueError):
        ctx.fit("se", rng.rand(10, 40), rng.rand(10), np.ones((1, 40)), [1.0], [1e-6], [0.0])   # d > 32
    # duplicated points and no noise: K is singular -> not positive definite
    Xd = np.vstack([X, X])
    with pytest.raises(np.linalg.LinAlgError):
        ctx.fit("se", Xd, rng.rand(20), [[0.3, 0.3]], [1.0], [0.0], [0.0])
    ctx.fit("se", X, rng.rand(10), [[0.3, 0.3]], [1.0], [1e-6], [0.0])        # the handle recovers
    with pytest.raises(ValueError):
        ctx.score(7, 0.0, X)                                       # unknown acquisition id
    with pytest.raises(ValueError):
        ctx.predict(rng.rand(5, 3))                                # wrong candidate dimension
    with pytest.raises(ValueError):
        ctx.set_precision(5, 1e-7)                                 # unknown precision path
    mu, s2 = ctx.predict(X)
    assert np.all(np.isfinite(mu)) and np.all(s2 > -1e-9)


def test_model_rejects_bad_input():
    from pybo_b200 import models
    gp = models.make_gp(1e-6, 1.0, [0.25, 0.25], 0.0)
    with pytest.raises(ValueError):
        gp.add_data(np.random.rand(3, 5), np.random.rand(3))       # wrong dimension
    with pytest.raises(ValueError):
        gp.add_data(np.random.rand(3, 2), np.random.rand(4))       # length mismatch
    with pytest.raises(ValueError):
        gp.get_improvement(0.0, np.random.rand(3, 2))              # acquisition without data
    with pytest.raises(ValueError):
        models.make_gp(1e-6, 1.0, [0.25], 0.0, kernel="periodic")
    with pytest.raises(ValueError):
        gp.set_precision("fp16")
