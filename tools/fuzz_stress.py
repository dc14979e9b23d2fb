"""Extended random-shape stress (GPU): tests/test_gpu_fuzz.py's case generator over more seeds than the test suite runs,
both precision paths against the oracle.  `BO_OZ_CHUNK_TILES=2 python tools/fuzz_stress.py --tier-min 64` shrinks the
candidate chunks to 256 and the tier-2 threshold to 64 so that the tiered levels of the int8 path (pilot chunk, mixed
levels, re-scoring one tier up) engage at the fuzz shapes' small candidate counts."""
import argparse
import sys

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import test_gpu_fuzz as F
from pybo_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--first", type=int, default=14)
ap.add_argument("--last", type=int, default=70)
ap.add_argument("--tier-min", type=int, default=0)
ap.add_argument("--tier-frac", type=float, default=-1.0)
a = ap.parse_args()
bad, tiers = [], {}
for seed in range(a.first, a.last):
    ctx = _lib.Context(0)
    if a.tier_min > 0:
        ctx.set_option("oz_tier_min", a.tier_min)
    if a.tier_frac >= 0:
        ctx.set_option("oz_tier_frac", a.tier_frac)
    try:
        F.test_random_shapes_both_paths(ctx, seed)
    except AssertionError as e:
        bad.append((seed, str(e)[:200]))
    except Exception as e:
        bad.append((seed, "EXC " + repr(e)[:200]))
    ctx.close()
print("seeds %d..%d failures: %s" % (a.first, a.last - 1, bad))
