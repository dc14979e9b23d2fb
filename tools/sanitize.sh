#!/bin/bash
# compute-sanitizer passes over tools/san_driver.py; logs land in gpurun_out/ (copied to profiles/ by hand)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  # BO_OZ_CHUNK_TILES=2: 256-candidate chunks, so the driver's 2^14 candidates exercise the tiered levels of the int8 path
  # BO_CHOL_FLOW_MIN=4: the persistent dataflow Cholesky (flag hand-over between resident CTAs) runs at the driver's small sizes too
  BO_OZ_CHUNK_TILES=2 BO_CHOL_FLOW_MIN=4 timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 python tools/san_driver.py > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/san_summary.txt
  tail -n 4 gpurun_out/san_$tool.log >> gpurun_out/san_summary.txt
done
