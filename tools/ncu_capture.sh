#!/bin/bash
# usage: ncu_capture.sh <name> <kernel regex> <launches to skip> <command...>
# One `ncu --set full` capture of the first matching launch after the skip; only the raw-metric CSV and the details
# page travel back (the .ncu-rep files are 20 MB each and gpurun merges at most 64 MiB).
name=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/$name "$@" > /dev/null 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
ls -la /tmp/$name.ncu-rep gpurun_out/${name}_raw.csv
