"""Dependent-issue latencies on the device (cycles): what bounds the serial pivot chain of the Cholesky."""
import sys
sys.path.insert(0, '.')
from pybo_b200 import _lib
ctx = _lib.Context(0)
for k in ("lat_dfma", "lat_dmma", "lat_rcp", "lat_rsqrt", "lat_syncthreads", "lat_mbarrier"):
    print("%-16s %.1f cycles" % (k, ctx.microbench(k, 4096)))
for k in ("dmma", "dfma", "dmma_dfma_mix"):
    print("%-16s %.1f TFLOP/s" % (k, ctx.microbench(k, 20000)))
