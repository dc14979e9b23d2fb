#!/bin/bash
# Tiered levels of the int8 path (main pass one half-level down, flagged list one level up, FP64 last): on / off per workload
mkdir -p gpurun_out
{
for wl in rbf_n4096_d8_ei matern_n4096_d8_ucb mixture32_n2048_d8_ei; do
  for t in 1; do
    echo "== $wl tiered=$t"
    timeout 600 python tools/level_ab.py --workload $wl --levels "" --tiered $t 2>&1 | tail -3
  done
done
} | tee gpurun_out/tier_ab.txt
timeout 1500 python -m pytest tests/test_gpu_rescue.py tests/test_gpu_ozaki.py tests/test_gpu_configs.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -15 | tee -a gpurun_out/tier_ab.txt
