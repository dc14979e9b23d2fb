for t in 0 1 2; do
  BO_OZ_HACK=$t python bench.py --steps 5 --warmup 3 --quick --cpu-seconds 1 > gpurun_out/hack_$t.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/hack_$t.json'))
k=d['kernels']
print("hack=$t ms/step %.2f clocks %s score avg %.4f ms" % (d['ms_per_step'], d['clocks']['sm_mhz'], k['oz_score_kernel']['ms']/k['oz_score_kernel']['launches']))
PY
done
