"""The int8-slice path and its FP64 rescue pass against the ORACLE (not against the FP64 GPU path):

* both precision paths on >= 2^16 Sobol candidates at every bench workload shape, held to the same 1e-6 with the
  SURVEY 8c(7) floor of 1e-12 max|ref| (config 2: the floor stated in test_gpu_configs.FLOOR2);
* the a-priori error bound behind the rescue decision really bounds the observed error of s2;
* what the rescue pass reports, and that a fit whose candidates mostly need FP64 is demoted to the FP64 path;
* regressions from round-1 review: scratch capacities across refits of different shape, UCB next to the data,
  top-k with fewer comparable values than k, precision as per-model state.
"""

import numpy as np
import pytest
from scipy.stats import qmc

from conftest import rel_err
from oracle import GPOracle, MixtureOracle, ucb_beta, ucb_index

pytestmark = pytest.mark.gpu
TOL = 1e-6


def problem(n, d, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    return rng, X, y, float(y.max() - y.min()), float(y.mean())


def _oracle_scores(gp, acq, param, Xc, batch=4096):
    out = []
    for i in range(0, len(Xc), batch):
        blk = Xc[i:i + batch]
        if acq == 1:
            out.append(gp.get_improvement(param, blk))
        elif acq == 2:
            out.append(gp.get_tail(param, blk))
        else:
            out.append(ucb_index(param, *gp.predict(blk)))
    return np.concatenate(out)


@pytest.mark.parametrize("name,kernel,n,d,acq,floor", [
    ("headline rbf_n4096_d8_ei", "se", 4096, 8, 1, 1e-12),
    ("config 3 matern_n4096_d8_ucb", "matern52", 4096, 8, 3, 1e-12),
    ("config 2 rbf_n1024_d4_ei", "se", 1024, 4, 1, 1e-9),
])
def test_bench_workloads_both_paths_vs_oracle_65536(ctx, name, kernel, n, d, acq, floor):
    """Same seeded problem as bench.py's `make_problem` (seed 0), first 2^16 points of the Sobol grid."""
    rng, X, y, rho, bias = problem(n, d, seed=0)
    gp = GPOracle(1e-6, rho, 0.25 * np.ones(d), bias, kernel)
    gp.add_data(X, y)
    ctx.fit(kernel, X, y, 0.25 * np.ones((1, d)), [rho], [1e-6], [bias])
    Xc = qmc.Sobol(d=d, scramble=False).random_base2(16)
    param = float(ucb_beta(n)) if acq == 3 else float(gp.predict(X)[0].max())
    ref = _oracle_scores(gp, acq, param, Xc)
    errs = {}
    for prec, label in ((0, "fp64"), (1, "int8")):
        ctx.set_precision(prec, 1e-8)
        val, _, best = ctx.score(acq, param, Xc, want_best=True)
        errs[label] = rel_err(val, ref, floor)
        assert errs[label] < TOL, (name, label, errs[label])
        assert best[1] == int(np.argmax(ref)), (name, label)
        idx, top = ctx.topk(10)
        assert np.array_equal(idx, np.lexsort((np.arange(len(ref)), -ref))[:10]), (name, label)
    ran8, rescued, total = ctx.rescue_info()
    print("%s: max rel err fp64 %.2e int8 %.2e (floor %g max|ref|), int8 path ran %s, %d of %d re-scored in FP64"
          % (name, errs["fp64"], errs["int8"], floor, ran8, rescued, total))


def test_bench_mixture_workload_both_paths_vs_oracle(ctx):
    """Config 5 (32 hyper-samples x n=2048, d=8, EI mixture): 2^13 candidates = 2^18 GP evaluations on the oracle."""
    rng, X, y, rho, bias = problem(2048, 8, seed=0)
    S = 32
    ell = np.tile(0.25 * np.ones(8), (S, 1)) * np.exp(0.1 * rng.randn(S, 8))
    rhos = np.full(S, rho) * np.exp(0.1 * rng.randn(S))
    sn2 = np.full(S, 1e-6) * np.exp(0.3 * rng.randn(S))
    biases = np.full(S, bias)
    gps = []
    for s in range(S):
        g = GPOracle(sn2[s], rhos[s], ell[s], biases[s], "se")
        g.add_data(X, y)
        gps.append(g)
    mix = MixtureOracle(gps)
    ctx.fit("se", X, y, ell, rhos, sn2, biases)
    Xc = qmc.Sobol(d=8, scramble=False).random_base2(13)
    target = float(mix.predict(X[:512])[0].max())
    ref = np.concatenate([mix.get_improvement(target, Xc[i:i + 2048]) for i in range(0, len(Xc), 2048)])
    for prec in (0, 1):
        ctx.set_precision(prec, 1e-8)
        val, _, best = ctx.score(1, target, Xc, want_best=True)
        assert rel_err(val, ref) < TOL, (prec, rel_err(val, ref))
        assert best[1] == int(np.argmax(ref))


@pytest.mark.parametrize("kernel,n,d,ellv,sn2,seed", [("se", 4096, 8, 0.25, 1e-6, 0), ("se", 1024, 4, 0.25, 1e-6, 0),
                                                     ("matern52", 383, 2, 0.3, 1e-4, 6), ("se", 1025, 12, 0.5, 1e-4, 7),
                                                     ("se", 200, 1, 0.3, 1e-4, 8)])
def test_error_bound_holds(ctx, kernel, n, d, ellv, sn2, seed):
    """|s2_int8 - s2_fp64| <= errK sqrt((rho - s2) rho) for every candidate, at several levels, including candidates
    placed on top of observations (where the kernel digits are largest); rescue off so the raw error is seen."""
    rng, X, y, rho, bias = problem(n, d, seed)
    ctx.fit(kernel, X, y, ellv * np.ones((1, d)), [rho], [sn2], [bias])
    Xc = np.concatenate([qmc.Sobol(d=d, scramble=False).random_base2(14),
                         np.clip(X[rng.randint(0, n, 2048)] + 1e-3 * rng.randn(2048, d), 0, 1)])
    mu0, s20 = ctx.predict(Xc)
    q = np.maximum(rho - s20, 0.0)
    ctx.set_rescue(False)
    for level in (4.0, 4.5, 5.0, 5.5):
        ctx.set_precision(1, level)
        errk = ctx.error_bound()[0]
        mu, s2 = ctx.predict(Xc)
        bound = errk * np.sqrt(q * rho)
        worst = float(np.max(np.abs(s2 - s20) / np.maximum(bound, 1e-300 + 1e-16 * rho)))
        assert worst < 1.0, (level, worst)
        assert np.max(np.abs(mu - mu0)) < 1e-9 * max(1.0, np.max(np.abs(mu0)))


def test_rescue_pass_reports_and_repairs(ctx):
    """A coarse level (3 slices) leaves large errors in s2; with the rescue pass on, the flagged candidates come back in
    FP64 and the result meets the tolerance against the oracle anyway; with it off the raw error is visible."""
    rng, X, y, rho, bias = problem(700, 6, seed=2)
    gp = GPOracle(1e-5, rho, 0.3 * np.ones(6), bias, "se")
    gp.add_data(X, y)
    ctx.fit("se", X, y, 0.3 * np.ones((1, 6)), [rho], [1e-5], [bias])
    Xc = qmc.Sobol(d=6, scramble=False).random_base2(15)[:40001]
    target = float(gp.predict(X)[0].max())
    ref = gp.get_improvement(target, Xc)
    ctx.set_precision(1, 4.0)
    ctx.set_rescue(True, 5e-7, 1e-12)
    val, _, best = ctx.score(1, target, Xc, want_best=True)
    ran8, rescued, total = ctx.rescue_info()
    assert ran8 and total == len(Xc) and 0 < rescued
    assert rel_err(val, ref) < TOL and best[1] == int(np.argmax(ref))
    ctx.set_rescue(False)
    raw, _, _ = ctx.score(1, target, Xc, want_best=True)
    assert ctx.rescue_info()[1] == 0
    assert rel_err(raw, ref) > rel_err(val, ref)
    # predict: s2 is what the rescue pass looks at
    ctx.set_rescue(True)
    mu, s2 = ctx.predict(Xc)
    rmu, rs2 = gp.predict(Xc)
    assert rel_err(mu, rmu, 1e-9) < TOL
    assert np.max(np.abs(s2 - rs2) / np.maximum(np.abs(rs2), 1e-9 * rho)) < TOL
    # at the training points everything is flagged (s2 ~ sn2): the incumbent target is an FP64 quantity
    ctx.set_rescue(True)                                          # (re-arm: a pass that rescued > 25 % demotes the fit to FP64)
    mu_d, s2_d = ctx.predict(X)
    assert ctx.rescue_info()[1] == len(X)
    assert rel_err(mu_d, gp.predict(X)[0], 1e-9) < TOL


def test_tiered_levels_headline_shape_vs_fp64_and_oracle(ctx):
    """Passes of >= 8 chunks run their first 4096 candidates one half-level below the level the tolerance selects and, if at
    most 10 % of it is flagged there, the rest too; the flagged list is re-scored on the int8 path at the selected
    level before FP64 takes what is left.  Every combination of tiers must meet the same 1e-6 against the FP64 path
    (and the oracle on a slice) with identical arg max and top-10: the tiers only decide the speed."""
    rng, X, y, rho, bias = problem(4096, 8, seed=0)
    gp = GPOracle(1e-6, rho, 0.25 * np.ones(8), bias, "se")
    gp.add_data(X, y)
    ctx.fit("se", X, y, 0.25 * np.ones((1, 8)), [rho], [1e-6], [bias])
    Xc = qmc.Sobol(d=8, scramble=False).random_base2(19)
    target = float(gp.predict(X)[0].max())
    ctx.set_precision(0, 1e-8)
    ref, _, rbest = ctx.score(1, target, Xc, want_best=True)
    rtop = ctx.topk(10)[0]
    sl = slice(40000, 40400)
    oref = gp.get_improvement(target, Xc[sl])
    ctx.set_precision(1, 1e-8)
    seen = {}
    for label, opts in (("untiered", dict(oz_tiered=0)),
                        ("tiered", dict(oz_tiered=1, oz_tier_frac=0.10, oz_tier_min=4096)),
                        ("pilot fails", dict(oz_tiered=1, oz_tier_frac=0.0, oz_tier_min=4096)),
                        ("straight to fp64", dict(oz_tiered=1, oz_tier_frac=0.10, oz_tier_min=1 << 30)),
                        ("tier 2 on short lists", dict(oz_tiered=1, oz_tier_frac=0.0, oz_tier_min=1))):
        for k, v in opts.items():
            ctx.set_option(k, v)
        ctx.set_rescue(True)                            # re-arm (a pass that hands > 25 % to FP64 demotes the fit)
        val, _, best = ctx.score(1, target, Xc, want_best=True)
        t = ctx.tier_info()
        seen[label] = t
        assert ctx.rescue_info()[0], label
        assert rel_err(val, ref) < 5e-7, (label, t)                # the rescue tolerance: what the path certifies
        assert rel_err(val[sl], oref) < TOL, (label, t)
        assert best[1] == rbest[1] and np.array_equal(ctx.topk(10)[0], rtop), (label, t)
    for k, v in dict(oz_tiered=1, oz_tier_frac=0.10, oz_tier_min=4096).items():
        ctx.set_option(k, v)
    assert seen["untiered"]["first"] == seen["untiered"]["rest"] == (5, False) and seen["untiered"]["tier2"] is None
    assert seen["tiered"]["first"] == (4, True) and seen["tiered"]["rest"] == (4, True), seen["tiered"]
    assert seen["tiered"]["tier2"] == (5, False) and seen["tiered"]["fp64_rescored"] < seen["tiered"]["first_flagged"]
    assert seen["pilot fails"]["first"] == (4, True) and seen["pilot fails"]["rest"] == (5, False)
    assert seen["straight to fp64"]["tier2"] is None
    assert seen["straight to fp64"]["fp64_rescored"] == seen["straight to fp64"]["first_flagged"] > 0
    assert seen["tier 2 on short lists"]["tier2"] == (5, True)
    # predict (s2 is what the tiers look at) on the same candidates
    mu, s2 = ctx.predict(Xc)
    assert ctx.tier_info()["first"] == (4, True)
    ctx.set_precision(0, 1e-8)
    rmu, rs2 = ctx.predict(Xc)
    assert rel_err(mu, rmu, 1e-9) < TOL
    assert np.max(np.abs(s2 - rs2) / np.maximum(np.abs(rs2), 1e-9 * rho)) < TOL


def test_refit_with_other_shapes_keeps_scratch_valid(ctx):
    """Round-1 review: gradient scratch was guarded by one capacity (np * cap).  Refit the same handle with other
    n, S and d and take gradients at batch sizes that make each buffer grow independently."""
    rng = np.random.RandomState(0)

    def check(n, d, S, M):
        X = rng.rand(n, d)
        y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
        rho, bias = float(np.ptp(y)) + 0.1, float(y.mean())
        ell = 0.3 * np.ones((S, d)) * np.exp(0.05 * rng.randn(S, d))
        rhos, sn2, biases = rho * np.ones(S), 1e-4 * np.ones(S), bias * np.ones(S)
        ctx.fit("se", X, y, ell, rhos, sn2, biases)
        gps = []
        for s in range(S):
            g = GPOracle(sn2[s], rhos[s], ell[s], biases[s], "se")
            g.add_data(X, y)
            gps.append(g)
        ref = gps[0] if S == 1 else MixtureOracle(gps)
        Xc = rng.rand(M, d)
        target = float(ref.predict(X)[0].max())
        val, grad, _ = ctx.score(1, target, Xc, grad=True)
        rv, rg = ref.get_improvement(target, Xc, grad=True)
        assert rel_err(val, rv, 1e-9) < TOL and rel_err(grad, rg, 1e-8) < 10 * TOL, (n, d, S, M)

    check(1024, 4, 1, 128)
    check(128, 4, 1, 1024)       # np * cap equal to the previous call's, Gpart / DmuS eight times larger
    check(128, 4, 10, 1024)      # S grows at the same n
    check(128, 16, 10, 1024)     # d grows
    check(600, 3, 2, 7)          # small-batch (GEMV) path after a large one
    check(64, 2, 3, 300)


def test_ucb_next_to_observations_is_finite(ctx):
    """rho - |v|^2 may round to <= 0 on top of an observation (sn2 = 1e-6 rho): UCB value and gradient stay finite,
    predict returns s2 >= 0, and top-k returns real indices."""
    rng, X, y, rho, bias = problem(500, 2, seed=4)
    ctx.fit("se", X, y, 0.25 * np.ones((1, 2)), [rho], [1e-8], [bias])
    Xc = np.concatenate([X[:300], X[:100] + 1e-9])
    for prec in (0, 1):
        ctx.set_precision(prec, 1e-8)
        val, _, best = ctx.score(3, 16.0, Xc, want_best=True)
        assert np.all(np.isfinite(val))
        mu, s2 = ctx.predict(Xc)
        assert np.all(s2 >= 0.0)
        idx, top = ctx.topk(10)
        assert len(idx) == 10 and np.all(idx >= 0) and np.all(idx < len(Xc))
    ctx.set_precision(0)
    v, g, _ = ctx.score(3, 16.0, Xc[:16], grad=True)
    assert np.all(np.isfinite(v)) and np.all(np.isfinite(g))


def test_topk_truncates_when_few_values_are_comparable(ctx):
    rng, X, y, rho, bias = problem(50, 2, seed=5)
    ctx.fit("se", X, y, 0.25 * np.ones((1, 2)), [rho], [1e-4], [bias])
    Xc = rng.rand(6, 2)
    ctx.score(1, 0.0, Xc)
    idx, val = ctx.topk(6)
    assert len(idx) == 6
    Xc[[1, 4]] = np.nan                                           # NaN candidates never rank
    v, _, best = ctx.score(0, 0.0, Xc, want_best=True)
    idx, val = ctx.topk(6)
    assert len(idx) == 4 and set(idx.tolist()) == {0, 2, 3, 5} and np.all(np.isfinite(val))
    assert best[1] in (0, 2, 3, 5)


def test_precision_is_per_model_state():
    """`model.copy()` shares the fitted handle (policies take one), but the precision path belongs to each model
    object: changing it on one does not change what the other computes."""
    from pybo_b200 import models
    rng, X, y, rho, bias = problem(300, 3, seed=6)
    a = models.make_gp(1e-4, rho, 0.3 * np.ones(3), bias)
    a.add_data(X, y)
    Xc = qmc.Sobol(d=3, scramble=False).random_base2(12)
    va = a.get_improvement(0.1, Xc)                                # fits lazily; copies taken from here on share the handle
    b = a.copy()
    assert b._fit is a._fit
    b.set_precision("int8", 3.0)                                  # deliberately coarse
    b._ensure_fit().set_rescue(False)
    vb = b.get_improvement(0.1, Xc)
    assert b._fit is a._fit and b._ensure_fit().rescue_info()[0]
    va2 = a.get_improvement(0.1, Xc)                              # a still scores in FP64 on the shared handle
    assert not a._ensure_fit().rescue_info()[0]
    assert np.array_equal(va, va2) and not np.array_equal(va, vb)
    a._ensure_fit().set_rescue(True)
    # ownership: dropping the copy releases its share at once (no reference cycle), so appends happen in place
    fit = a._fit
    assert fit.owners == 2
    del b
    assert fit.owners == 1
    a.add_data(rng.rand(3), 0.5)
    assert a._fit is fit and fit.ctx.n == 301
