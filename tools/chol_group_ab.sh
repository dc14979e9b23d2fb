#!/bin/bash
# A/B of the batched Cholesky: one stream pair vs two half-batches on two stream pairs (BO_CHOL_LANES), and the
# panel-group size of the trailing update (BO_CHOL_GROUP).
mkdir -p gpurun_out
for cfg in "BO_CHOL_LANES=2" "BO_CHOL_LANES=4" "BO_CHOL_LANES=8" "BO_CHOL_LANES=4 BO_CHOL_GROUP=2" "BO_CHOL_LANES=4 BO_CHOL_GROUP=3" "BO_CHOL_LANES=8 BO_CHOL_GROUP=2"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/chol_bench.py 2>&1 | grep -A8 "batch=32"
done | tee gpurun_out/chol_group_ab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_mcmc.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/chol_group_ab.txt
