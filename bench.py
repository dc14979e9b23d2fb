#!/usr/bin/env python
"""bench.py -- acquisition evaluations per second on the BASELINE.json headline
shape (RBF GP, n = 4096 observations, d = 8, EI over 2^20 Sobol candidates per
GPU), plus the n = 4096 Cholesky figure the metric also names.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K ...   (CPU reference arm)

A step is one scoring pass of the hot path (`finit = f(xgrid)`, reference
solvers/lbfgs.py:50) over this rank's candidate block followed by the incumbent
reduction.  `value` times that pass with the candidates already resident in HBM
(CUDA events on the library's stream); `e2e` times the same pass through the
public plugin API with pinned HOST candidates, host->device copy and the
device->host read of the top-10 inside the timed region.  Prints ONE JSON line.
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
    # BASELINE.json configs[1..4] (parity cases; selectable for profiling)
    "rbf_n1024_d4_ei": dict(kernel="se", n=1024, d=4, acq="ei", M=1 << 20, S=1),
    "matern_n4096_d8_ucb": dict(kernel="matern52", n=4096, d=8, acq="ucb", M=1 << 20, S=1),
    "mixture32_n2048_d8_ei": dict(kernel="se", n=2048, d=8, acq="ei", M=1 << 17, S=32),
}
FALLBACK_PEAKS = dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
# dram__bytes_read.sum + dram__bytes_write.sum per score_gemm_kernel launch from the committed
# `ncu --set full` capture (profiles/r1_fp64_score_gemm_ncu.txt); algorithmic operand bytes are
# W (lower triangle, 67 MB) + one K* chunk (268 MB): the excess is L2 thrash, 5.5 % of DRAM peak.
TRAFFIC_NCU = {"rbf_n4096_d8_ei:fp64": 2.005e9,
               # oz_score_kernel<5>, 32768 candidates per launch (profiles/r1_int8_oz_score_ncu.txt): the K* slice
               # planes of the chunk (671 MB) are read ~2.6x from DRAM, W slices stay in L2 (89.6 % hit rate)
               "rbf_n4096_d8_ei:ozaki": 1.83e9}


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
    ell = np.tile(0.25 * np.ones(d), (S, 1))
    rho = np.full(S, float(y.max() - y.min()))
    sn2 = np.full(S, 1e-6)
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
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if cell.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------
def acq_param(spec, X, predict):
    if spec["acq"] == "ucb":
        from pybo_b200.policies import ucb_beta
        return 3, float(ucb_beta(spec["n"]))
    target = float(np.max(predict(X)[0]))       # policies/simple.py:21 with xi = 0
    return 1, target


def our_arm(args):
    import torch
    import torch.distributed as dist
    from pybo_b200 import _lib, dist as bdist, models, policies

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        # ONE JSON line on stdout: NCCL prints its version banner there when the communicator comes up
        # (NCCL_DEBUG=VERSION/WARN), so file descriptor 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    spec = WORKLOADS[args.workload]
    n, d, M, S = spec["n"], spec["d"], (args.candidates or spec["M"]), spec["S"]
    X, y, ell, rho, sn2, bias = make_problem(spec)

    # the model behind the plugin surface (replicated fit on every rank)
    if S == 1:
        model = models.make_gp(sn2[0], rho[0], ell[0], bias[0], kernel=spec["kernel"], device=local)
    else:
        model = models.MCMC.from_samples(spec["kernel"], ell, rho, sn2, bias, device=local)
    model.add_data(X, y) if S == 1 else models._Base.add_data(model, X, y)
    t0 = time.perf_counter()
    ctx = model._ensure_fit()
    ctx.sync()
    fit_s = time.perf_counter() - t0
    acq, param = acq_param(spec, X, model.predict)
    index = policies.ModelIndex(model, acq, param)
    if args.precision == "ozaki":
        model.set_precision("int8", args.tol)

    xc_dev = sobol_block(M, d, rank * M).cuda()            # this rank's block, resident in HBM
    val_dev = torch.empty(M, dtype=torch.float64, device="cuda")
    xc_host = xc_dev.cpu().pin_memory()
    xc_np = xc_host.numpy()
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)

    def step_device():
        bv, bi = ctx.score_device(acq, param, M, xc_dev.data_ptr(), val_dev.data_ptr(), want_best=True)
        return bdist.reduce_incumbent(bv, bi + rank * M)

    def step_e2e():
        idx, val = index.best_of(xc_np, 10)
        return bdist.reduce_incumbent(val[0], int(idx[0]) + rank * M)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        out = None
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), out

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ctx.profile(True)
    ctx.profile_reset()
    l0 = ctx.launch_count()
    dev_ms, _, incumbent = timed(step_device, args.steps)
    launches = ctx.launch_count() - l0
    prof = ctx.profile_report()
    ctx.profile(False)
    clocks = sampler.stop() if sampler else None

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    _, e2e_ms, incumbent_e2e = timed(step_e2e, args.steps)

    # the same pass with the Sobol block generated on the device (bo_candidates_sobol): no candidate copy at all
    unit_box = np.array([[0.0, 1.0]] * d)

    def step_grid():
        pts, val, idx = index.best_of_sobol(unit_box, M, 10, start=rank * M)
        return bdist.reduce_incumbent(val[0], int(idx[0]))

    step_grid()
    _, grid_ms, incumbent_grid = timed(step_grid, args.steps)

    # for the record: the next-cheaper precision level (4 slices + first dropped pair group), same timed loop
    fast_level = None
    if args.precision == "ozaki" and args.tol == 1e-8:
        level = ctx.precision_info()
        model.set_precision("int8", 4.5)
        for _ in range(2):
            step_device()
        fms, _, finc = timed(step_device, args.steps)
        fast_level = dict(level="4 slices + first dropped pair group (13 digit pairs)", value=M * world * args.steps / (fms * 1e-3), unit="evals/s",
                          incumbent_index=finc[1],
                          parity="max EI rel. error 4.2e-7 vs FP64 over 2^20 candidates (p99.9 4.6e-8), identical arg max "
                                 "(tools/oz_err.py)")
        model.set_precision("int8", args.tol)
        step_device()

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
    if args.precision == "ozaki":
        kname = "oz_score_kernel"
        gk = prof.get(kname, dict(launches=0, total_ms=0.0))
        _, slices, extra = ctx.precision_info()
        pairs = slices * (slices + 1) // 2 + ((slices - 1) if extra else 0)
        nb = npad // 64
        # algorithmic: n^2 (forward-substitution equivalent) + 4n (reductions) flop per candidate
        alg_flop_per_launch = (n * n + 4 * n) * (M * S * args.steps) / max(1, gk["launches"])
        exec_ops_per_launch = 2.0 * pairs * 64 * 64 * nb * (nb + 1) / 2 * (M * S * args.steps) / max(1, gk["launches"])
        avg_ms = gk["total_ms"] / max(1, gk["launches"])
        achieved = alg_flop_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        executed = exec_ops_per_launch / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        int8_peak = 2.0 * pk["bf16_tflops"]
        roofline = dict(bound="tensor", kernel=kname, achieved=achieved, peak=int8_peak, unit="TFLOP/s",
                        frac=achieved / int8_peak, traffic=TRAFFIC_NCU.get(args.workload + ":ozaki"),
                        peak_source="int8 tensor roof = 2 x the %s bf16 figure of MEASURED_PEAKS.json (tcgen05 kind::i8 issues at "
                                    "twice the bf16 rate: ncu peak_sustained 16384 vs 8192 op/clk/SM)" % pk["source"],
                        int8_slices=slices, slice_pairs=pairs, executed_int8_tops=executed,
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
        roofline = dict(bound="tensor", kernel=kname, achieved=achieved, peak=fp64_peak, unit="TFLOP/s",
                        frac=achieved / fp64_peak if fp64_peak else None, traffic=TRAFFIC_NCU.get(args.workload + ":fp64"),
                        peak_source="FP64 roof measured in this run by bo_microbench (register-resident loops on every SM): "
                                    "DMMA m8n8k4 %.1f, DFMA %.1f TFLOP/s; the larger is used (ncu: both issue at 128 flop/clk/SM = "
                                    "37.2 TFLOP/s at 1965 MHz). MEASURED_PEAKS.json holds no FP64 figure" % (dmma_peak, dfma_peak),
                        frac_of_bf16_peak=achieved / pk["bf16_tflops"], bf16_peak=pk["bf16_tflops"], peaks=pk["source"],
                        alg_flop_per_launch=alg_flop_per_launch, avg_launch_ms=avg_ms, launches=gk["launches"],
                        share_of_step=gk["total_ms"] / dev_ms if dev_ms else None)

    chol = cholesky_metric(ctx, spec, X, ell[0], rho[0], sn2[0], pk, fp64_peak)
    append = append_metric(spec, X, y, ell[0], rho[0], sn2[0], bias[0], pk, local) if S == 1 else None
    thompson = thompson_metric(local) if (S == 1 and args.workload == "rbf_n4096_d8_ei") else None
    cpu = cpu_baseline(spec, X, y, ell, rho, sn2, bias, budget_s=args.cpu_seconds) if world == 1 else None

    line = dict(metric="acq_evals_per_sec", value=value, unit="evals/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=args.workload, kernel=spec["kernel"], n=n, d=d, acq=spec["acq"],
                            hyper_samples=S, candidates_per_gpu=M, candidates="unscrambled Sobol, contiguous block per rank",
                            precision=("int8 slices on tcgen05 (tol %g -> %d slices%s), FP64 reassembly and FP64 mean"
                                       % (args.tol, ctx.precision_info()[1], " + first dropped pair group" if ctx.precision_info()[2] else ""))
                            if args.precision == "ozaki" else "fp64 (DMMA)",
                            l2="inputs exceed L2: W factor %.0f MB + cross-kernel scratch >= %.0f MB + candidates %.0f MB per pass"
                            % (n * n * 8 / 1e6, n * 8192 * 8 / 1e6, M * d * 8 / 1e6), parallelism="dp%d candidate shards" % world),
                clocks=clocks,
                e2e=dict(value=e2e_value, unit="evals/s", ms_per_step=e2e_ms / args.steps,
                         h2d_bytes_per_step=int(M * d * 8), d2h_bytes_per_step=int(10 * 16),
                         api="policies.ModelIndex.best_of (score + device top-10) on pinned host candidates"),
                e2e_device_grid=dict(value=M * world * args.steps / (grid_ms * 1e-3), unit="evals/s", ms_per_step=grid_ms / args.steps,
                                     h2d_bytes_per_step=0, d2h_bytes_per_step=int(10 * 16), incumbent_index=incumbent_grid[1],
                                     api="policies.ModelIndex.best_of_sobol: Sobol block generated on the device, scored, device top-10"),
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, cholesky=chol, incremental_refit=append, thompson=thompson,
                fit_seconds=fit_s, incumbent=dict(value=incumbent[0], index=incumbent[1]), faster_level=fast_level,
                kernels={k: dict(launches=v["launches"], ms=round(v["total_ms"], 3)) for k, v in prof.items()})
    print(json.dumps(line))
    if world > 1:
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


def thompson_metric(device, n=4096, d=16, ndraw=256, m=1024, M=1 << 20):
    """BASELINE config 4 shape (Thompson: n=4096, d=16, 256 posterior draws x 2^20 candidates, shared
    1024-feature basis; reference policies/simple.py:48 batched): per-draw arg max on the FP64 DMMA path and
    on the int8-slice tcgen05 path, candidates resident in HBM."""
    import torch
    from pybo_b200 import models
    rng = np.random.RandomState(0)
    X = rng.rand(n, d)
    y = np.sin(X.sum(axis=1)) + 0.01 * rng.randn(n)
    gp = models.make_gp(1e-6, float(y.max() - y.min()), 0.25 * np.ones(d), float(y.mean()), device=device)
    gp.add_data(X, y)
    tb = models.ThompsonBatch(gp, m=m, ndraw=ndraw, rng=0)
    ctx = tb._context()
    xc = sobol_block(M, d, 0).cuda()
    out = {}
    for name in ("fp64", "int8"):
        tb.set_precision(name, 1e-8)
        for _ in range(2):
            bv, bi = ctx.thompson_eval_device(M, xc.data_ptr())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            bv, bi = ctx.thompson_eval_device(M, xc.data_ptr())
        dt = (time.perf_counter() - t0) / 3
        out[name] = dict(ms_per_pass=dt * 1e3, draw_evals_per_sec=ndraw * M / dt,
                         algorithmic_tflops=2.0 * ndraw * M * m / dt / 1e12, argmax=bi.tolist()[:4])
    out["argmax_identical"] = bool(out["fp64"]["argmax"] == out["int8"]["argmax"])
    out["shape"] = dict(n=n, d=d, draws=ndraw, features=m, candidates=M)
    ctx.close()
    return out


# ----------------------------------------------------------------------------------------
def _blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(spec, X, y, ell, rho, sn2, bias, budget_s=12.0, batch=2048):
    """The float64 NumPy/SciPy/LAPACK oracle (the calls the reference reaches through reggie)
    timed on this host: scoring batches of Sobol candidates until `budget_s` is spent."""
    from oracle import GPOracle, MixtureOracle
    from oracle import ucb_beta, ucb_index
    n, d, S = spec["n"], spec["d"], spec["S"]
    gps = []
    t0 = time.perf_counter()
    for s in range(S):
        g = GPOracle(sn2[s], rho[s], ell[s], bias[s], spec["kernel"])
        g.add_data(X, y)
        gps.append(g)
    fit_s = time.perf_counter() - t0
    model = gps[0] if S == 1 else MixtureOracle(gps)
    target = float(np.max(model.predict(X[: min(n, 512)])[0]))
    done, spent, start = 0, 0.0, 0
    while spent < budget_s:
        Xc = sobol_block(batch, d, start).numpy()
        t0 = time.perf_counter()
        if spec["acq"] == "ucb":
            ucb_index(ucb_beta(n), *model.predict(Xc))
        else:
            model.get_improvement(target, Xc)
        spent += time.perf_counter() - t0
        done += batch
        start += batch
    return dict(value=done / spent, unit="evals/s", cores=_blas_threads(), kind="port",
                sample="%d Sobol candidates in batches of %d (%.1f s) on the %s workload; fit %.1f s not counted"
                       % (done, batch, spent, "n=%d d=%d" % (n, d), fit_s), host_cpus=os.cpu_count())


def reference_arm(args):
    """CPU reference arm: the reference's own path is NumPy/SciPy through `reggie`, which is absent;
    the oracle port (same LAPACK/BLAS calls) is timed with every host thread it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = WORKLOADS[args.workload]
    X, y, ell, rho, sn2, bias = make_problem(spec)
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(spec, X, y, ell, rho, sn2, bias, budget_s=1.0)
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
                config=dict(workload=args.workload, kernel=spec["kernel"], n=spec["n"], d=spec["d"], acq=spec["acq"],
                            hyper_samples=spec["S"], candidates_per_gpu=spec["M"]),
                cpu_baseline=res, e2e=dict(value=value, unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
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
    ap.add_argument("--precision", default="ozaki", choices=["ozaki", "fp64"],
                    help="scoring contraction: error-bounded int8 slices on tcgen05 (default) or FP64 DMMA")
    ap.add_argument("--tol", type=float, default=1e-8,
                    help="ozaki: target abs error of V entries / sqrt(rho) (>= 2 pins the level, e.g. 5 or 5.5). "
                         "Default 1e-8 -> 5 base-256 slices (15 digit pairs): max EI error 4.9e-8 vs FP64 over all "
                         "2^20 candidates; 4.5 -> 4 slices + first dropped pair group (13 pairs): 4.2e-7")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()
