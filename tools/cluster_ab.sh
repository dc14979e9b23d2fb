for c in 2 4; do
  echo "== tests with BO_OZ_CLUSTER=$c"
  BO_OZ_CLUSTER=$c timeout 400 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_rescue.py -m gpu -q -x 2>&1 | tail -3
done
for c in 1 2 4; do
  BO_OZ_CLUSTER=$c timeout 200 python bench.py --steps 5 --warmup 3 --quick --cpu-seconds 1 > gpurun_out/cl_$c.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/cl_$c.json'))
    k=d['kernels']
    print("cluster=$c value %.4g ms/step %.2f clocks %s score avg %.4f ms parity %.2e %s" % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], k['oz_score_kernel']['ms']/k['oz_score_kernel']['launches'], d['parity_in_run']['max_rel_err'], d['parity_in_run']['passed']))
except Exception as e:
    print("cluster=$c failed", e)
PY
done
