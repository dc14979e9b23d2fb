timeout 300 python -m pytest tests/test_gpu_ozaki.py -m gpu -q -x 2>&1 | tail -2
for c in 256 512 1024; do
  BO_OZ_CHUNK_TILES=$c timeout 200 python bench.py --steps 8 --warmup 3 --quick --cpu-seconds 1 > gpurun_out/ch_$c.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/ch_$c.json'))
k=d['kernels']
print("chunk_tiles=$c value %.4g ms/step %.2f clocks %s score/cand-tile %.5f slicer/cand-tile %.5f parity %s" % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], k['oz_score_kernel']['ms']/8/8192*1e3, k['oz_kstar_slices_kernel']['ms']/8/8192*1e3, d['parity_in_run']['passed']))
PY
done
