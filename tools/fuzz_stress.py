import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import test_gpu_fuzz as F
from pybo_b200 import _lib
bad = []
for seed in range(14, 70):
    ctx = _lib.Context(0)
    try:
        F.test_random_shapes_both_paths(ctx, seed)
    except AssertionError as e:
        bad.append((seed, str(e)[:200]))
    except Exception as e:
        bad.append((seed, "EXC " + repr(e)[:200]))
    ctx.close()
print("failures:", bad)
