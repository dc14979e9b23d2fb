#!/bin/bash
# compute-sanitizer passes over tools/san_driver.py; logs land in gpurun_out/ (copied to profiles/ by hand)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 python tools/san_driver.py > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/san_summary.txt
  tail -n 4 gpurun_out/san_$tool.log >> gpurun_out/san_summary.txt
done
