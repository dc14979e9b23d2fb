#!/bin/bash
# Round-2 profiling pass after the tiered levels went in (1 GPU): bench record, launch list, full capture of the main-level contraction.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --quick --cpu-seconds 1 > gpurun_out/b_ncu.log 2>&1
bash tools/ncu_capture.sh prof_oz_score oz_score 40 python bench.py --steps 1 --warmup 3 --quick --cpu-seconds 1 --candidates 524288
tail -c 300 gpurun_out/bench_r2b.err
