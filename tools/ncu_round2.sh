#!/bin/bash
# Round-2 profiling pass (1 GPU): launch list of a short headline run, full captures of the dominant kernels, SASS opcode summary.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --quick --cpu-seconds 1 > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:oz_score -s 40 -c 1 -o gpurun_out/prof_oz_score python bench.py --steps 1 --warmup 3 --quick --cpu-seconds 1 --candidates 65536 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:oz_kstar -s 40 -c 1 -o gpurun_out/prof_oz_kstar python bench.py --steps 1 --warmup 3 --quick --cpu-seconds 1 --candidates 65536 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:chol_flow -c 1 -o gpurun_out/prof_chol_flow python tools/chol_bench.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_tile -c 1 -o gpurun_out/prof_gram python tools/fit_bench.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
