"""Slice-level A/B on the bench workloads (GPU): for each pinned level (S, extra) time the device-resident scoring pass,
report how many candidates the FP64 rescue pass re-scored and the in-run parity against the FP64 path."""
import argparse
import sys
import time

sys.path.insert(0, '.')
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="rbf_n4096_d8_ei")
    ap.add_argument("--levels", default="4.0,4.5,5.0")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--tiered", type=int, default=1)
    ap.add_argument("--rescue-tol", type=float, default=0.0)
    a = ap.parse_args()
    args = argparse.Namespace(gpus=1)
    h = bench.Harness(args)
    w = bench.ScoringWorkload(h, a.workload)
    w.ctx.set_option("oz_tiered", a.tiered)
    if a.rescue_tol > 0:
        w.ctx.set_rescue(True, a.rescue_tol, 1e-12)
    for lv in [float(x) for x in a.levels.split(",") if x] + [1e-8]:
        w.set_path("ozaki", lv)
        for _ in range(2):
            w.step_device()
        ms, wall, _ = h.timed(w.ctx, w.step_device, a.steps)
        lvl = w.level()
        par = w.parity()
        print("level %-6g  %.2f ms/step  %.3e evals/s  %s  max_rel_err=%.2e argmax=%s top10=%s passed=%s" % (
            lv, ms / a.steps, w.M * a.steps / (ms * 1e-3), lvl, par["max_rel_err"], par["argmax_identical"],
            par["top10_identical"], par["passed"]), flush=True)


if __name__ == "__main__":
    main()
