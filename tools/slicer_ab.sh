for t in 0 1; do
  BO_OZ_EXP2_TABLE=$t python bench.py --steps 5 --warmup 3 --quick --cpu-seconds 1 > gpurun_out/ab_$t.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_$t.json'))
k=d['kernels']
print("table=$t value %.4g ms/step %.2f clocks %s slicer avg %.4f ms score avg %.4f ms parity %.2e" % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], k['oz_kstar_slices_kernel']['ms']/k['oz_kstar_slices_kernel']['launches'], k['oz_score_kernel']['ms']/k['oz_score_kernel']['launches'], d['parity_in_run']['max_rel_err']))
PY
done
