#!/usr/bin/env python
"""bench.py -- acquisition evaluations per second on the BASELINE.json headline
shape (RBF GP, n = 4096 observations, d = 8, EI over 2^20 Sobol candidates per
GPU), plus the n = 4096 Cholesky figure the metric also names.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K ...   (CPU reference arm)

A step is one scoring pass of the hot path (`finit = f(xgrid)`, reference
solvers/lbfgs.py:50) over this rank's candidate block followed by the incumbent
exchange.  `value` times that pass with the candidates already resident in HBM
(CUDA events on the library's stream); `e2e` times the same pass through the
public plugin API with pinned HOST candidates, host->device copy and the
device->host read of the top-10 inside the timed region.  The default run also
scores a few steps of every other BASELINE config (`configs`) and checks the
int8-slice path against the FP64 path inside the run (`parity_in_run`).
Prints ONE JSON line.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json metric: "acq evals/sec (n=4096 GP, d=8)"; SURVEY 8d headline
    "rbf_n4096_d8_ei": dict(kernel="se", n=4096, d=8, acq="ei", M=1 << 20, S=1),
    # BASELINE.json configs[1..4]
    "rbf_n1024_d4_ei": dict(kernel="se", n=1024, d=4, acq="ei", M=1 << 20, S=1),
    "matern_n4096_d8_ucb": dict(kernel="matern52", n=4096, d=8, acq="ucb", M=1 << 20, S=1),
    "thompson_n4096_d16": dict(kernel="se", n=4096, d=16, acq="thompson", M=1 << 20, S=1, ndraw=256, m=1024,
                               ell=0.25, sn2=1e-6),
    "mixture32_n2048_d8_ei": dict(kernel="se", n=2048, d=8, acq="ei", M=1 << 17, S=32),
}
CONFIG_NAMES = {"rbf_n1024_d4_ei": "config 2", "matern_n4096_d8_ucb": "config 3", "thompson_n4096_d16": "config 4",
                "mixture32_n2048_d8_ei": "config 5"}
FALLBACK_PEAKS = dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
PARITY_TOL = 1e-6          # north_star: outputs within 1e-6 relative
PARITY_FLOOR = 1e-12       # ... with the SURVEY 8c(7) floor of 1e-12 max|ref|


def workload_config(name, spec, M, world):
    """The `config` object both arms print (identical keys and values, so the driver can match the two lines)."""
    n, d = spec["n"], spec["d"]
    cfg = dict(workload=name, kernel=spec["kernel"], n=n, d=d, acq=spec["acq"], hyper_samples=spec["S"], candidates_per_gpu=int(M),
               candidates="unscrambled Sobol, contiguous block per rank",
               l2="inputs exceed L2: W factor %.0f MB + cross-kernel scratch >= %.0f MB + candidates %.0f MB per pass"
                  % (n * n * 8 / 1e6, n * 8192 * 8 / 1e6, M * d * 8 / 1e6),
               parallelism="dp%d candidate shards" % world)
    if spec["acq"] == "thompson":
        cfg.update(draws=spec["ndraw"], features=spec["m"], basis="shared by all draws")
    return cfg


def flop_per_eval(n, d):
    """F(n,d) = n^2 + n(3d+2) + 4n + 30 flop per acquisition evaluation."""
    return n * n + n * (3 * d + 2) + 4 * n + 30


def make_problem(spec, seed=0):
    """SURVEY 8d synthetic inputs: X ~ U[0,1]^(n x d), y = sin(sum x) + 0.01 N(0,1), hypers in the
    reference's default regime (bayesopt.py:98-102): ell = width/4, rho = range(y), sn2 = 1e-6."""
    rng = np.random.RandomState(seed)
    n, d, S = spec["n"], spec["d"], spec["S"]
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    ell = np.tile(spec.get("ell", 0.25) * np.ones(d), (S, 1))
    rho = np.full(S, float(y.max() - y.min()))
    sn2 = np.full(S, spec.get("sn2", 1e-6))
    bias = np.full(S, float(y.mean()))
    if S > 1:                                   # log-normally jittered hyper-samples, seed 0
        ell = ell * np.exp(0.1 * rng.randn(S, d))
        rho = rho * np.exp(0.1 * rng.randn(S))
        sn2 = sn2 * np.exp(0.3 * rng.randn(S))
    return X, y, ell, rho, sn2, bias


def sobol_block(M, d, start):
    """Candidates [start, start + M) of the unscrambled Sobol sequence in [0,1]^d."""
    import torch
    eng = torch.quasirandom.SobolEngine(d, scramble=False)
    if start:
        eng.fast_forward(start)
    return eng.draw(M, dtype=torch.float64)


def peaks():
    p = dict(FALLBACK_PEAKS, source="fallback")
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            m = json.load(fh)
        p.update(hbm_gbs=float(m["hbm_gbs"]), bf16_tflops=float(m["bf16_tflops"]),
                 bf16_tflops_sustained=float(m.get("bf16_tflops_sustained", m["bf16_tflops"])), source="measured")
    except Exception:
        pass
    return p


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed summary of an
    `ncu --set full` capture (profiles/traffic.json, written by tools/profile_summary.py); None when no capture
    of this workload / kernel is on file."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)
        e = t.get(workload, {}).get(kernel)
        return dict(bytes=float(e["dram_bytes"]), source=e.get("source")) if e else None
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        threading.Thread.__init__(self, daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if cell.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w=float(np.median(pw)) if pw else None, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------
def acq_param(spec, X, predict):
    if spec["acq"] == "ucb":
        from pybo_b200.policies import ucb_beta
        return 3, float(ucb_beta(spec["n"]))
    target = float(np.max(predict(X)[0]))       # policies/simple.py:21 with xi = 0
    return 1, target


class Harness(object):
    """Process-group plumbing and the timed loop shared by every workload of a run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world != args.gpus:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            # ONE JSON line on stdout: NCCL prints its version banner there when the communicator comes up
            # (NCCL_DEBUG=VERSION/WARN), so file descriptor 1 points at stderr until the first collective is done
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
                warm = torch.zeros(1, device="cuda")
                dist.all_reduce(warm)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, ctx, fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events on the library's stream, max over ranks.
        Returns (device ms, wall ms, last result)."""
        torch = self.torch
        stream = torch.cuda.ExternalStream(ctx.stream, device=self.local)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        self.barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), out

    def all_true(self, flag):
        t = self.torch.tensor([1.0 if flag else 0.0], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def sum(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t)
        return float(t.item())

    def max(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


class ScoringWorkload(object):
    """One GP scoring workload (headline or configs 2, 3, 5) on this rank's candidate block."""

    def __init__(self, h, name, M=None, precision="ozaki", tol=1e-8):
        from pybo_b200 import models, policies
        self.h, self.name, self.spec = h, name, WORKLOADS[name]
        spec = self.spec
        self.M = int(M or spec["M"])
        self.n, self.d, self.S = spec["n"], spec["d"], spec["S"]
        self.X, self.y, self.ell, self.rho, self.sn2, self.bias = make_problem(spec)
        if self.S == 1:
            self.model = models.make_gp(self.sn2[0], self.rho[0], self.ell[0], self.bias[0], kernel=spec["kernel"], device=h.local)
            self.model.add_data(self.X, self.y)
        else:
            self.model = models.MCMC.from_samples(spec["kernel"], self.ell, self.rho, self.sn2, self.bias, device=h.local)
            models._Base.add_data(self.model, self.X, self.y)
        t0 = time.perf_counter()
        self.ctx = self.model._ensure_fit()
        self.ctx.sync()
        self.fit_s = time.perf_counter() - t0
        self.acq, self.param = acq_param(spec, self.X, self.model.predict)
        self.index = policies.ModelIndex(self.model, self.acq, self.param)
        self.tol = tol
        self.set_path(precision)
        torch = h.torch
        self.xc_dev = sobol_block(self.M, self.d, h.rank * self.M).cuda()       # this rank's block, resident in HBM
        self.val_dev = torch.empty(self.M, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()

    def set_path(self, precision, tol=None):
        tol = self.tol if tol is None else tol
        self.model.set_precision("int8" if precision == "ozaki" else "fp64", tol)
        self.ctx = self.model._ensure_fit()                                      # writes the path into the handle
        self.precision = precision

    def step_device(self):
        """Scoring pass over the resident block, then the incumbent exchange: the (value, index) record stays on
        the device, is all-gathered over NCCL on the library's stream and merged by one kernel."""
        from pybo_b200 import dist as bdist
        rec = self.ctx.score_incumbent(self.acq, self.param, self.M, self.xc_dev.data_ptr(), offset=self.h.rank * self.M,
                                       val_ptr=self.val_dev.data_ptr())
        v, i = bdist.exchange_incumbents(self.ctx, rec, 1)
        return float(v[0]), int(i[0])

    def level(self):
        """What the last pass actually ran: path, slice level(s), tiers of the repair."""
        ran8, rescued, total = self.ctx.rescue_info()
        if not ran8:
            return dict(path="fp64", slices=None, extra_group=None, digit_pairs=None, rescued_fp64=0, candidates=int(total))
        t = self.ctx.tier_info()
        pairs = lambda lv: lv[0] * (lv[0] + 1) // 2 + ((lv[0] - 1) if lv[1] else 0)
        slices, extra = t["rest"]
        out = dict(path="int8 slices", slices=slices, extra_group=extra, digit_pairs=pairs(t["rest"]),
                   rescued_fp64=int(rescued), candidates=int(total), flagged_by_main_pass=int(t["first_flagged"]))
        if t["first"] != t["rest"]:
            out["first_chunk"] = dict(slices=t["first"][0], extra_group=t["first"][1], digit_pairs=pairs(t["first"]))
        if t["tier2"]:
            out["flagged_rescored_at"] = dict(slices=t["tier2"][0], extra_group=t["tier2"][1], digit_pairs=pairs(t["tier2"]))
        return out

    def parity(self, m_check=None):
        """In-run parity of the selected path against the FP64 path on the first m_check candidates of this rank's
        block (default: all of them): max relative error with the floor, identical arg max and top-10."""
        torch = self.h.torch
        m = int(min(self.M, m_check or self.M))
        ptr = self.xc_dev.data_ptr()
        keep = self.precision
        v = torch.empty(m, dtype=torch.float64, device="cuda")
        bv, bi = self.ctx.score_device(self.acq, self.param, m, ptr, v.data_ptr(), want_best=True)
        lvl = self.level()
        top = self.ctx.topk(10)[0]
        self.set_path("fp64")
        ref = torch.empty(m, dtype=torch.float64, device="cuda")
        rv, ri = self.ctx.score_device(self.acq, self.param, m, ptr, ref.data_ptr(), want_best=True)
        rtop = self.ctx.topk(10)[0]
        self.set_path(keep)
        scale = float(ref.abs().max())
        scale = self.h.max(scale)
        err = float(((v - ref).abs() / torch.clamp(ref.abs(), min=PARITY_FLOOR * scale)).max())
        err = self.h.max(err)
        same_arg = self.h.all_true(bi == ri)
        same_top = self.h.all_true(bool(np.array_equal(top, rtop)))
        return dict(against="FP64 (DMMA) path of the same library, same candidates", candidates_per_gpu=m,
                    max_rel_err=err, floor="%g * max|ref|" % PARITY_FLOOR, tol=PARITY_TOL,
                    argmax_identical=same_arg, top10_identical=same_top,
                    rescued_fp64=int(self.h.sum(lvl["rescued_fp64"])), passed=bool(err < PARITY_TOL and same_arg and same_top))


class ThompsonWorkload(object):
    """BASELINE config 4: 256 posterior draws x 2^20 candidates per GPU at n = 4096, d = 16; draws built on the
    device from the same seed on every rank, per-draw arg max exchanged as 256 packed records."""

    def __init__(self, h, name="thompson_n4096_d16", M=None, shared_basis=True, m=None):
        from pybo_b200 import dist as bdist, models
        self.h, self.name, self.spec = h, name, WORKLOADS[name]
        spec = self.spec
        self.M = int(M or spec["M"])
        self.n, self.d, self.ndraw = spec["n"], spec["d"], spec["ndraw"]
        self.m = int(m or spec["m"])
        self.shared = shared_basis
        X, y, ell, rho, sn2, bias = make_problem(spec)
        gp = models.make_gp(sn2[0], rho[0], ell[0], bias[0], kernel=spec["kernel"], device=h.local)
        gp.add_data(X, y)                                # data only (fits are lazy): the draw construction needs no GP factor
        h.torch.cuda.synchronize()
        t0 = time.perf_counter()
        self.tb = models.ThompsonBatch(gp, m=self.m, ndraw=self.ndraw, rng=0, shared_basis=shared_basis)
        self.ctx = self.tb._context()
        self.ctx.sync()
        self.build_s = time.perf_counter() - t0
        self.sharded = bdist.ShardedThompson(self.tb, rank=h.rank, world=h.world)
        self.xc_dev = sobol_block(self.M, self.d, h.rank * self.M).cuda()
        h.torch.cuda.synchronize()

    def set_path(self, path):
        self.tb.set_precision(path, 1e-8)
        self.ctx = self.tb._context()

    def step_device(self):
        return self.sharded.argmax_device(self.M, self.xc_dev.data_ptr(), self.h.rank * self.M)


def bench_config(h, name, steps, precision, tol):
    """A few device-timed steps of one BASELINE config, sharded like the headline, with in-run parity."""
    if WORKLOADS[name]["acq"] == "thompson":
        w = ThompsonWorkload(h, name)
        out = {}
        results = {}
        for path in ("fp64", "int8"):
            w.set_path(path)
            w.step_device()
            ms, _, res = h.timed(w.ctx, w.step_device, steps)
            results[path] = res
            out[path] = dict(value=w.ndraw * w.M * h.world * steps / (ms * 1e-3), ms_per_step=ms / steps)
        same = bool(np.array_equal(results["fp64"][1], results["int8"][1]))
        entry = dict(config=CONFIG_NAMES[name], metric="draw_evals_per_sec", unit="draw-evals/s", value=out["int8"]["value"],
                     ms_per_step=out["int8"]["ms_per_step"], steps=steps, path="int8 slices (cosine features x Theta on tcgen05)",
                     fp64_path=out["fp64"], draws=w.ndraw, features=w.m, basis="shared", candidates_per_gpu=w.M,
                     build_seconds=w.build_s, build="bo_thompson_build: features, Phi^T Phi + sn2 I, Cholesky, solves on the device",
                     parity_in_run=dict(against="FP64 path (DMMA contraction, on-the-fly cosine features)",
                                        argmax_identical_all_draws=h.all_true(same), passed=h.all_true(same)))
        w.ctx.close()
        return entry
    w = ScoringWorkload(h, name, precision=precision, tol=tol)
    for _ in range(2):                                    # (an int8 pass that rescues > 25 % demotes the fit to FP64:
        w.step_device()                                   #  the timed steps run the path the library settled on)
    ms, _, inc = h.timed(w.ctx, w.step_device, steps)
    lvl = w.level()
    par = w.parity(m_check=1 << 17) if lvl["path"] != "fp64" else dict(
        against="n/a: the int8 path handed this fit to the FP64 path (its rescue pass had to re-score more than a "
                "quarter of the candidates)", passed=True)
    entry = dict(config=CONFIG_NAMES[name], metric="acq_evals_per_sec", unit="evals/s", value=w.M * h.world * steps / (ms * 1e-3),
                 ms_per_step=ms / steps, steps=steps, level=lvl, hyper_samples=w.S, candidates_per_gpu=w.M,
                 incumbent_index=inc[1], parity_in_run=par)
    w.ctx.close()
    return entry


def our_arm(args):
    h = Harness(args)
    if WORKLOADS[args.workload]["acq"] == "thompson":
        return thompson_arm(h, args)
    torch, dist = h.torch, h.dist
    from pybo_b200 import dist as bdist
    rank, local, world = h.rank, h.local, h.world
    w = ScoringWorkload(h, args.workload, M=args.candidates or None, precision=args.precision, tol=args.tol)
    spec, n, d, M, S, ctx, index = w.spec, w.n, w.d, w.M, w.S, w.ctx, w.index
    xc_host = w.xc_dev.cpu().pin_memory()
    xc_np = xc_host.numpy()

    def step_e2e():
        idx, val = index.best_of(xc_np, 10)
        return bdist.reduce_incumbent(val[0], int(idx[0]) + rank * M)

    for _ in range(args.warmup):
        w.step_device()
    ctx = w.ctx
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ctx.profile(True)
    ctx.profile_reset()
    l0 = ctx.launch_count()
    dev_ms, _, incumbent = h.timed(ctx, w.step_device, args.steps)
    launches = ctx.launch_count() - l0
    prof = ctx.profile_report()
    ctx.profile(False)
    clocks = sampler.stop() if sampler else None
    level = w.level()

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    _, e2e_ms, incumbent_e2e = h.timed(ctx, step_e2e, args.steps)

    # the same pass with the Sobol block generated on the device (bo_candidates_sobol): no candidate copy at all
    unit_box = np.array([[0.0, 1.0]] * d)

    def step_grid():
        pts, val, idx = index.best_of_sobol(unit_box, M, 10, start=rank * M)
        return bdist.reduce_incumbent(val[0], int(idx[0]))

    step_grid()
    _, grid_ms, incumbent_grid = h.timed(ctx, step_grid, max(2, args.steps // 4))
    grid_steps = max(2, args.steps // 4)

    # in-run parity of the timed path against the FP64 path over ALL candidates, and the FP64 path's own speed
    parity = w.parity() if args.precision == "ozaki" else None
    fp64_path = None
    if args.precision == "ozaki" and not args.quick:
        w.set_path("fp64")
        w.step_device()
        fsteps = max(1, min(3, args.steps))
        fms, _, finc = h.timed(w.ctx, w.step_device, fsteps)
        fp64_path = dict(value=M * world * fsteps / (fms * 1e-3), unit="evals/s", ms_per_step=fms / fsteps, steps=fsteps,
                         incumbent_index=finc[1])
        w.set_path("ozaki")
        w.step_device()

    # for the record: the same pass with the tiers off (every chunk at the level the tolerance selects, flagged
    # candidates straight to FP64) -- the round-1 / early round-2 configuration
    fast_level = None
    if args.precision == "ozaki" and not args.quick:
        w.ctx.set_option("oz_tiered", 0)
        for _ in range(2):
            w.step_device()
        fsteps = max(2, args.steps // 4)
        fms, _, finc = h.timed(w.ctx, w.step_device, fsteps)
        fl = w.level()
        fpar = w.parity(m_check=1 << 18)
        fast_level = dict(level=fl, value=M * world * fsteps / (fms * 1e-3), unit="evals/s", ms_per_step=fms / fsteps,
                          incumbent_index=finc[1], parity_in_run=fpar,
                          note="tiers off (bo_set_option oz_tiered = 0): one level for the whole pass, flagged candidates "
                               "straight to the FP64 path")
        w.ctx.set_option("oz_tiered", 1)
        w.step_device()

    configs = None
    if args.workload == "rbf_n4096_d8_ei" and not args.quick and not args.candidates:
        # free the headline's device buffers the other configs do not need
        configs = {}
        for name in ("rbf_n1024_d4_ei", "matern_n4096_d8_ucb", "mixture32_n2048_d8_ei", "thompson_n4096_d16"):
            configs[name] = bench_config(h, name, args.config_steps, args.precision, args.tol)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = M * world * args.steps / (dev_ms * 1e-3)
    e2e_value = M * world * args.steps / (e2e_ms * 1e-3)

    # roofline of the dominant kernel: the triangular contraction V = W K* (+ fused reductions)
    dmma_peak = ctx.microbench("dmma")
    dfma_peak = ctx.microbench("dfma")
    fp64_peak = max(dmma_peak, dfma_peak)
    npad = -(-n // 128) * 128
    if level["path"] != "fp64":
        kname = "oz_score_kernel"
        gk = prof.get(kname, dict(launches=0, total_ms=0.0))
        slices, extra, pairs = level["slices"], level["extra_group"], level["digit_pairs"]
        nb = npad // 64
        # algorithmic: n^2 (forward-substitution equivalent) + 4n (reductions) flop per candidate
        alg_flop_per_launch = (n * n + 4 * n) * (M * S * args.steps) / max(1, gk["launches"])
        # executed int8 products: the main pass at its level(s) plus the flagged list re-scored one tier up
        first_len = 4096 if "first_chunk" in level else 0          # the pilot chunk of a tiered pass that kept the selected level
        first_pairs = level.get("first_chunk", {}).get("digit_pairs", pairs)
        t2 = level.get("flagged_rescored_at")
        pair_cands = pairs * max(0, M - first_len) + first_pairs * min(M, first_len) + (t2["digit_pairs"] * level["flagged_by_main_pass"] if t2 else 0)
        exec_ops_per_launch = 2.0 * 64 * 64 * nb * (nb + 1) / 2 * (pair_cands * S * args.steps) / max(1, gk["launches"])
        avg_ms = gk["total_ms"] / max(1, gk["launches"])
        achieved = alg_flop_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        executed = exec_ops_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        int8_peak = 2.0 * pk["bf16_tflops"]
        tr = ncu_traffic(args.workload, kname)
        roofline = dict(bound="tensor", kernel=kname, achieved=achieved, peak=int8_peak, unit="TFLOP/s",
                        frac=achieved / int8_peak, traffic=tr["bytes"] if tr else None, traffic_source=tr["source"] if tr else None,
                        peak_source="int8 tensor roof = 2 x the %s bf16 figure of MEASURED_PEAKS.json (tcgen05 kind::i8 issues at "
                                    "twice the bf16 rate: ncu peak_sustained 16384 vs 8192 op/clk/SM)" % pk["source"],
                        int8_slices=slices, slice_pairs=pairs, rescored_one_tier_up=level.get("flagged_by_main_pass") if t2 else 0,
                        executed_int8_tops=executed,
                        frac_executed=executed / int8_peak,
                        note="achieved counts ALGORITHMIC flop (n^2 + 4n per candidate); the emulation executes %d int8 "
                             "products per algorithmic product, so frac_executed is the tensor-pipe utilisation" % pairs,
                        fp64_roof=fp64_peak, achieved_over_fp64_roof=achieved / fp64_peak if fp64_peak else None,
                        alg_flop_per_launch=alg_flop_per_launch, avg_launch_ms=avg_ms, launches=gk["launches"],
                        share_of_step=gk["total_ms"] / dev_ms if dev_ms else None)
    else:
        kname = "score_gemm_kernel"
        gk = prof.get(kname, dict(launches=0, total_ms=0.0))
        alg_flop_per_launch = (n * n + 4 * n) * (M * S * args.steps) / max(1, gk["launches"])
        avg_ms = gk["total_ms"] / max(1, gk["launches"])
        achieved = alg_flop_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        tr = ncu_traffic(args.workload, kname)
        roofline = dict(bound="tensor", kernel=kname, achieved=achieved, peak=fp64_peak, unit="TFLOP/s",
                        frac=achieved / fp64_peak if fp64_peak else None, traffic=tr["bytes"] if tr else None,
                        traffic_source=tr["source"] if tr else None,
                        peak_source="FP64 roof measured in this run by bo_microbench (register-resident loops on every SM): "
                                    "DMMA m8n8k4 %.1f, DFMA %.1f TFLOP/s; the larger is used (ncu: both issue at 128 flop/clk/SM = "
                                    "37.2 TFLOP/s at 1965 MHz). MEASURED_PEAKS.json holds no FP64 figure" % (dmma_peak, dfma_peak),
                        frac_of_bf16_peak=achieved / pk["bf16_tflops"], bf16_peak=pk["bf16_tflops"], peaks=pk["source"],
                        alg_flop_per_launch=alg_flop_per_launch, avg_launch_ms=avg_ms, launches=gk["launches"],
                        share_of_step=gk["total_ms"] / dev_ms if dev_ms else None)

    quick = args.quick
    chol = cholesky_metric(ctx, spec, w.X, w.ell[0], w.rho[0], w.sn2[0], pk, fp64_peak)
    chol_b = None if quick else cholesky_batched_metric(ctx, pk, fp64_peak)
    fit = None if quick else fit_metric(ctx, spec, w.X, pk)
    append = append_metric(spec, w.X, w.y, w.ell[0], w.rho[0], w.sn2[0], w.bias[0], pk, local) if (S == 1 and not quick) else None
    cpu = cpu_baseline(spec, w.X, w.y, w.ell, w.rho, w.sn2, w.bias, budget_s=args.cpu_seconds) if world == 1 else None

    if level["path"] != "fp64":
        dtype = ("int8 slices: %d balanced base-256 digits per operand%s, exact int32 accumulation on tcgen05, f64 reassembly; "
                 "f64 mean; candidates whose a-priori error bound exceeds 5e-7 re-scored%s in f64"
                 % (level["slices"], " + first dropped pair group" if level["extra_group"] else "",
                    " with %d digits, what that cannot certify" % level["flagged_rescored_at"]["slices"] if level.get("flagged_rescored_at") else ""))
    else:
        dtype = "f64"
    line = dict(metric="acq_evals_per_sec", value=value, unit="evals/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype=dtype, data="synthetic",
                config=workload_config(args.workload, spec, M, world),
                path=dict(precision=level,
                          incumbent_exchange="packed device record -> NCCL all-gather on the library's stream -> device merge"
                          if world > 1 else "device record -> merge kernel -> one 16-byte read-back"),
                clocks=clocks,
                e2e=dict(value=e2e_value, unit="evals/s", ms_per_step=e2e_ms / args.steps,
                         h2d_bytes_per_step=int(M * d * 8), d2h_bytes_per_step=int(10 * 16),
                         api="policies.ModelIndex.best_of (score + device top-10) on pinned host candidates"),
                e2e_device_grid=dict(value=M * world * grid_steps / (grid_ms * 1e-3), unit="evals/s", ms_per_step=grid_ms / grid_steps,
                                     h2d_bytes_per_step=0, d2h_bytes_per_step=int(10 * 16), incumbent_index=incumbent_grid[1],
                                     api="policies.ModelIndex.best_of_sobol: Sobol block generated on the device, scored, device top-10"),
                gpu_launches=int(launches), roofline=roofline, parity_in_run=parity, fp64_path=fp64_path, cpu_baseline=cpu,
                cholesky=chol, cholesky_batched=chol_b, fit=fit, incremental_refit=append, configs=configs,
                fit_seconds=w.fit_s, incumbent=dict(value=incumbent[0], index=incumbent[1]), untiered=fast_level,
                kernels={k: dict(launches=v["launches"], ms=round(v["total_ms"], 3)) for k, v in prof.items()})
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def thompson_arm(h, args):
    """`--workload thompson_n4096_d16` (BASELINE config 4) as the timed workload: draw evaluations per second."""
    dist = h.dist
    w = ThompsonWorkload(h, args.workload, M=args.candidates or None)
    path = "int8" if args.precision == "ozaki" else "fp64"
    w.set_path(path)
    for _ in range(args.warmup):
        w.step_device()
    sampler = ClockSampler(h.local) if h.rank == 0 else None
    if sampler:
        sampler.start()
    w.ctx.profile(True)
    w.ctx.profile_reset()
    l0 = w.ctx.launch_count()
    ms, _, res = h.timed(w.ctx, w.step_device, args.steps)
    launches = w.ctx.launch_count() - l0
    prof = w.ctx.profile_report()
    w.ctx.profile(False)
    clocks = sampler.stop() if sampler else None
    # end to end through the public API: pinned host candidates in, 256 (value, index) pairs out
    xc_np = w.xc_dev.cpu().pin_memory().numpy()
    lo = h.rank * w.M

    def step_e2e():
        rec, nd = w.ctx.thompson_incumbents(w.M, xc_np, offset=lo, flags=0)
        from pybo_b200 import dist as bdist
        return bdist.exchange_incumbents(w.ctx, rec, nd)

    step_e2e()
    _, e2e_ms, res_e2e = h.timed(w.ctx, step_e2e, args.steps)
    # the other path, for parity in the run
    other = "fp64" if path == "int8" else "int8"
    w.set_path(other)
    w.step_device()
    oms, _, ores = h.timed(w.ctx, w.step_device, max(1, min(3, args.steps)))
    osteps = max(1, min(3, args.steps))
    same = h.all_true(bool(np.array_equal(res[1], ores[1])))
    # one basis per draw = 256 independent sample_f calls (m = 128 features each), the literal reading of config 4
    per_draw = None
    if not args.quick:
        wp = ThompsonWorkload(h, args.workload, M=min(w.M, 1 << 17), shared_basis=False, m=128)
        wp.step_device()
        pms, _, _ = h.timed(wp.ctx, wp.step_device, 2)
        per_draw = dict(value=wp.ndraw * wp.M * h.world * 2 / (pms * 1e-3), unit="draw-evals/s", ms_per_step=pms / 2,
                        candidates_per_gpu=wp.M, features=128, build_seconds=wp.build_s,
                        note="every draw has its own 128-feature basis (256 feature systems built and factored on the device); "
                             "evaluation is the SIMT kernel (no shared contraction to put on the tensor cores)")
        wp.ctx.close()
    if h.rank != 0:
        if h.world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    value = w.ndraw * w.M * h.world * args.steps / (ms * 1e-3)
    kname = "oz_thompson_kernel" if path == "int8" else "thompson_gemm_kernel"
    gk = prof.get(kname, dict(launches=0, total_ms=0.0))
    alg = 2.0 * w.ndraw * w.m * w.M * args.steps / max(1, gk["launches"])
    avg_ms = gk["total_ms"] / max(1, gk["launches"])
    achieved = alg / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
    peak = 2.0 * pk["bf16_tflops"] if path == "int8" else max(w.ctx.microbench("dmma"), w.ctx.microbench("dfma"))
    cpu = thompson_cpu_baseline(w, budget_s=args.cpu_seconds) if h.world == 1 else None
    line = dict(metric="thompson_draw_evals_per_sec", value=value, unit="draw-evals/s", n_gpus=h.world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="int8 slices of cosine features and Theta on tcgen05, f64 reassembly" if path == "int8" else "f64",
                data="synthetic",
                config=workload_config(args.workload, w.spec, w.M, h.world),
                path=dict(draws_built="on the device from the same seed on every rank (bo_thompson_build, %.3f s)" % w.build_s),
                clocks=clocks,
                e2e=dict(value=w.ndraw * w.M * h.world * args.steps / (e2e_ms * 1e-3), unit="draw-evals/s", ms_per_step=e2e_ms / args.steps,
                         h2d_bytes_per_step=int(w.M * w.d * 8), d2h_bytes_per_step=int(w.ndraw * 16)),
                gpu_launches=int(launches),
                roofline=dict(bound="tensor", kernel=kname, achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak if peak else None,
                              traffic=None, alg_flop_per_launch=alg, avg_launch_ms=avg_ms, launches=gk["launches"],
                              note="algorithmic 2 m flop per draw-evaluation; the int8 path executes S(S+1)/2 products per algorithmic one"),
                parity_in_run=dict(against="%s path" % other, argmax_identical_all_draws=same, passed=same,
                                   other_path=dict(value=w.ndraw * w.M * h.world * osteps / (oms * 1e-3), ms_per_step=oms / osteps)),
                per_draw_basis=per_draw, cpu_baseline=cpu,
                kernels={k: dict(launches=v["launches"], ms=round(v["total_ms"], 3)) for k, v in prof.items()})
    print(json.dumps(line))
    if h.world > 1:
        dist.destroy_process_group()


def cholesky_metric(ctx, spec, X, ell, rho, sn2, pk, fp64_peak):
    """The metric's second half: n x n Cholesky, algorithmic bytes n(n+1)*8 over its time."""
    import torch
    n = spec["n"]
    K = torch.from_numpy(ctx.gram(spec["kernel"], X, ell, rho, sn2)).cuda()
    work = torch.empty_like(K)
    best = None
    for rep in range(5):
        work.copy_(K)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.cholesky_device(n, 1, work.data_ptr())
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    gbs = n * (n + 1) * 8 / best / 1e9
    tfl = n ** 3 / 3.0 / best / 1e12
    return dict(n=n, ms=best * 1e3, algorithmic_bytes=n * (n + 1) * 8, gbs=gbs, frac_hbm=gbs / pk["hbm_gbs"],
                hbm_peak_gbs=pk["hbm_gbs"], tflops=tfl, frac_fp64=tfl / fp64_peak if fp64_peak else None,
                note="compute-bound: n^3/3 flop over n(n+1)*8 bytes = %.0f flop/B; the HBM fraction cannot approach 1 in fp64" % (n / 24.0))


def cholesky_batched_metric(ctx, pk, fp64_peak, n=2048, batch=32):
    """Config 5's factorisation: 32 independent n = 2048 Cholesky factorisations in one batched call."""
    import torch
    rng = np.random.RandomState(1)
    X = rng.rand(n, 8)
    K1 = torch.from_numpy(ctx.gram("se", X, 0.25 * np.ones(8), 1.0, 1e-6)).cuda()
    K = K1.unsqueeze(0).repeat(batch, 1, 1).contiguous()
    K += 1e-6 * torch.arange(batch, device="cuda", dtype=torch.float64).view(-1, 1, 1) * torch.eye(n, device="cuda", dtype=torch.float64)
    work = torch.empty_like(K)
    best = None
    for rep in range(3):
        work.copy_(K)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.cholesky_device(n, batch, work.data_ptr())
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    nbytes = batch * n * (n + 1) * 8
    tfl = batch * n ** 3 / 3.0 / best / 1e12
    return dict(n=n, batch=batch, ms=best * 1e3, algorithmic_bytes=nbytes, gbs=nbytes / best / 1e9, frac_hbm=nbytes / best / 1e9 / pk["hbm_gbs"],
                tflops=tfl, frac_fp64=tfl / fp64_peak if fp64_peak else None)


def fit_metric(ctx, spec, X, pk):
    """Gram kernel alone (CUDA events through the library's profiler): n^2 * 8 bytes written (lower triangle only
    inside bo_fit: n(n+1)/2 * 8)."""
    from pybo_b200 import _lib
    n, d = spec["n"], spec["d"]
    c = _lib.Context(ctx.device)
    rng = np.random.RandomState(0)
    y = rng.randn(n)
    c.fit(spec["kernel"], X, y, 0.25 * np.ones((1, d)), [1.0], [1e-6], [0.0])
    c.profile(True)
    c.profile_reset()
    for _ in range(3):
        c.fit(spec["kernel"], X, y, 0.25 * np.ones((1, d)), [1.0], [1e-6], [0.0])
    c.sync()
    prof = c.profile_report()
    c.profile(False)
    c.close()
    out = {}
    for k, v in prof.items():
        out[k] = dict(launches=v["launches"], ms_per_fit=v["total_ms"] / 3)
    g = prof.get("gram_kernel")
    if g and g["launches"]:
        us = 1e3 * g["total_ms"] / g["launches"]
        nbytes = n * (n + 64) / 2 * 8                       # lower-triangular 64-tiles
        out["gram_summary"] = dict(us=us, bytes_written=nbytes, gbs=nbytes / (us * 1e-6) / 1e9,
                                   frac_hbm=nbytes / (us * 1e-6) / 1e9 / pk["hbm_gbs"])
    return out


def append_metric(spec, X, y, ell, rho, sn2, bias, pk, device, k=32):
    """Incremental refit (bo_append, reference bayesopt.py:269 `model.add_data`): the last k observations
    appended one at a time to a fit of the first n - k.  Algorithmic bytes per append: the lower triangle of
    W and the upper triangle of W^T streamed once each = n^2 * 8 B; HBM-bound."""
    from pybo_b200 import _lib
    n = spec["n"]
    c = _lib.Context(device)
    c.fit(spec["kernel"], X[:n - k], y[:n - k], ell[None], [rho], [sn2], [bias])
    c.append(X[n - k], y[n - k:n - k + 1])
    c.sync()
    t0 = time.perf_counter()
    for i in range(n - k + 1, n):
        c.append(X[i], y[i:i + 1])
    c.sync()
    dt = (time.perf_counter() - t0) / (k - 1)
    # the two matrix-vector kernels alone (CUDA events): the last 8 appends again on a fresh fit
    c.fit(spec["kernel"], X[:n - 8], y[:n - 8], ell[None], [rho], [sn2], [bias])
    c.profile(True)
    c.profile_reset()
    for i in range(n - 8, n):
        c.append(X[i], y[i:i + 1])
    c.sync()
    prof = c.profile_report()
    c.profile(False)
    kern = {}
    for name in ("append_wk_kernel", "append_wrow_kernel"):
        if name in prof and prof[name]["launches"]:
            us = 1e3 * prof[name]["total_ms"] / prof[name]["launches"]
            kgbs = (n - 4) * (n - 4) * 4 / (us * 1e-6) / 1e9        # one triangle of doubles
            kern[name] = dict(us=us, gbs=kgbs, frac_hbm=kgbs / pk["hbm_gbs"])
    c.sync()
    t0 = time.perf_counter()
    c.fit(spec["kernel"], X, y, ell[None], [rho], [sn2], [bias])
    c.sync()
    refit = time.perf_counter() - t0
    c.close()
    gbs = n * n * 8 / dt / 1e9
    return dict(n=n, ms_per_append=dt * 1e3, full_refit_ms=refit * 1e3, algorithmic_bytes=n * n * 8, gbs=gbs,
                frac_hbm=gbs / pk["hbm_gbs"], kernels=kern,
                note="ms_per_append / gbs: wall clock per bo_append call incl. its host sync (4 kernels); "
                     "kernels: the two triangular matrix-vector products alone, CUDA events")


# ----------------------------------------------------------------------------------------
def _host_threads():
    """Threads the CPU arm may use: every core of the box.  torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would silently halve-or-worse the LAPACK baseline, so the BLAS pools are set explicitly."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except (AttributeError, OSError):
        pass
    return n


def cpu_baseline(spec, X, y, ell, rho, sn2, bias, budget_s=12.0, batch=2048):
    """The float64 NumPy/SciPy/LAPACK oracle (the calls the reference reaches through reggie) timed on this host
    with every core: scoring batches of Sobol candidates until `budget_s` is spent.  Two forms of the same algebra
    are timed, half the budget each: `fast` (scaled squared distance through one dgemm: the fastest honest
    NumPy/SciPy form, reported as `value`) and `broadcast` (the (M, n, d) difference tensor the checker uses,
    round 1's figure)."""
    from threadpoolctl import threadpool_limits
    from oracle import GPOracle, MixtureOracle, predict_fast
    from oracle import ucb_beta, ucb_index, ei_from_moments
    n, d, S = spec["n"], spec["d"], spec["S"]
    threads = _host_threads()
    with threadpool_limits(limits=threads):
        gps = []
        t0 = time.perf_counter()
        for s in range(S):
            g = GPOracle(sn2[s], rho[s], ell[s], bias[s], spec["kernel"])
            g.add_data(X, y)
            gps.append(g)
        fit_s = time.perf_counter() - t0
        model = gps[0] if S == 1 else MixtureOracle(gps)
        target = float(np.max(model.predict(X[: min(n, 512)])[0]))
        beta = ucb_beta(n)

        def score_fast(Xc):
            if spec["acq"] == "ucb":
                return ucb_index(beta, *predict_fast(gps[0], Xc))
            return np.mean([ei_from_moments(target, *predict_fast(g, Xc)) for g in gps], axis=0)

        def score_broadcast(Xc):
            if spec["acq"] == "ucb":
                return ucb_index(beta, *model.predict(Xc))
            return model.get_improvement(target, Xc)

        out = {}
        start = 0
        for label, fn in (("fast", score_fast), ("broadcast", score_broadcast)):
            done, spent = 0, 0.0
            while spent < 0.5 * budget_s:
                Xc = sobol_block(batch, d, start).numpy()
                t0 = time.perf_counter()
                fn(Xc)
                spent += time.perf_counter() - t0
                done += batch
                start += batch
            out[label] = (done, spent)
    done, spent = out["fast"]
    return dict(value=done / spent, unit="evals/s", cores=threads, kind="port",
                sample="%d Sobol candidates in batches of %d (%.1f s) on the %s workload, dgemm-based distance; fit %.1f s not counted"
                       % (done, batch, spent, "n=%d d=%d" % (n, d), fit_s), host_cpus=os.cpu_count(),
                broadcast_distance=dict(value=out["broadcast"][0] / out["broadcast"][1], unit="evals/s",
                                        sample="%d candidates (%.1f s), (M, n, d) difference tensor as in round 1"
                                               % out["broadcast"]),
                omp_num_threads_env=os.environ.get("OMP_NUM_THREADS"))


def thompson_cpu_baseline(w, budget_s=12.0, batch=4096):
    """NumPy/BLAS form of the Thompson batch evaluation, F = bias + scale cos(X W^T + b) Theta^T, and per-draw arg max."""
    from threadpoolctl import threadpool_limits
    threads = _host_threads()
    tb = w.tb
    W, b, theta = tb.W[0], tb.b[0], tb.theta
    done, spent, start = 0, 0.0, 0
    with threadpool_limits(limits=threads):
        while spent < budget_s:
            Xc = sobol_block(batch, w.d, start).numpy()
            t0 = time.perf_counter()
            F = tb.bias + (tb.scale * np.cos(Xc @ W.T + b)) @ theta.T
            F.argmax(axis=0)
            spent += time.perf_counter() - t0
            done += batch
            start += batch
    return dict(value=w.ndraw * done / spent, unit="draw-evals/s", cores=threads, kind="port",
                sample="%d candidates x %d draws in batches of %d (%.1f s)" % (done, w.ndraw, batch, spent))


def reference_arm(args):
    """CPU reference arm: the reference's own path is NumPy/SciPy through `reggie`, which is absent;
    the oracle port (same LAPACK/BLAS calls) is timed with every host thread it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = WORKLOADS[args.workload]
    if spec["acq"] == "thompson":
        return reference_arm_thompson(args, spec)
    X, y, ell, rho, sn2, bias = make_problem(spec)
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(spec, X, y, ell, rho, sn2, bias, budget_s=2.0)
    vals, t0 = [], time.perf_counter()
    res = None
    for _ in range(args.steps):
        res = cpu_baseline(spec, X, y, ell, rho, sn2, bias, budget_s=per_step)
        vals.append(res["value"])
    value = float(np.mean(vals))
    res["value"] = value
    line = dict(impl="reference", metric="acq_evals_per_sec", value=value, unit="evals/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=(time.perf_counter() - t0) * 1e3 / max(1, args.steps),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=workload_config(args.workload, spec, args.candidates or spec["M"], args.gpus),
                cpu_baseline=res, e2e=dict(value=value, unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


def reference_arm_thompson(args, spec):
    from oracle import GPOracle, thompson_batch_oracle
    X, y, ell, rho, sn2, bias = make_problem(spec)

    class _W(object):
        pass
    w = _W()
    w.d, w.ndraw = spec["d"], spec["ndraw"]
    gp = GPOracle(sn2[0], rho[0], ell[0], bias[0], spec["kernel"])
    gp.X, gp.Y = X, y                                       # data only; the draw construction needs no GP factor
    t0 = time.perf_counter()
    Wm, b, theta, scale = thompson_batch_oracle(gp, spec["m"], spec["ndraw"], rng=0)
    build_s = time.perf_counter() - t0

    class _TB(object):
        pass
    w.tb = _TB()
    w.tb.W, w.tb.b, w.tb.theta, w.tb.scale, w.tb.bias = Wm, b, theta, scale, float(bias[0])
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    vals, t0 = [], time.perf_counter()
    res = None
    for _ in range(args.steps):
        res = thompson_cpu_baseline(w, budget_s=per_step)
        vals.append(res["value"])
    value = float(np.mean(vals))
    res["value"] = value
    res["build_seconds"] = build_s
    line = dict(impl="reference", metric="thompson_draw_evals_per_sec", value=value, unit="draw-evals/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=(time.perf_counter() - t0) * 1e3 / max(1, args.steps), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=workload_config(args.workload, spec, args.candidates or spec["M"], args.gpus),
                cpu_baseline=res, e2e=dict(value=value, unit="draw-evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rbf_n4096_d8_ei", choices=sorted(WORKLOADS))
    ap.add_argument("--candidates", type=int, default=0, help="override candidates per GPU (profiling only)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--config-steps", type=int, default=3, help="timed steps of each other BASELINE config in `configs`")
    ap.add_argument("--quick", action="store_true", help="headline numbers only (no configs / fp64 / untiered legs)")
    ap.add_argument("--precision", default="ozaki", choices=["ozaki", "fp64"],
                    help="scoring contraction: error-bounded int8 slices on tcgen05 (default) or FP64 DMMA")
    ap.add_argument("--tol", type=float, default=1e-8,
                    help="ozaki: target abs error of V entries / sqrt(rho) (>= 2 pins the level, e.g. 5 or 5.5). "
                         "Default 1e-8 selects 5 base-256 slices (15 digit pairs) at the headline shape; the tiers then run the main "
                         "pass at 4 slices + extra (13 pairs) and re-score what it flags at 5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
