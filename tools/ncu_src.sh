#!/bin/bash
# tools/ncu_src.sh <tag> <kernel-regex> <skip> <python args...>: ncu --set full with source correlation; raw + SASS + CUDA-line CSV -> gpurun_out/
tag=$1; shift; rex=$1; shift; skip=$1; shift
timeout 400 ncu --set full --import-source on --clock-control none -k regex:$rex -s $skip -c 1 -f -o /tmp/ncu_$tag python "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_${tag}_src.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page source --print-source cuda --csv > gpurun_out/ncu_${tag}_cuda.csv 2>/dev/null
ls -la gpurun_out/ncu_${tag}_raw.csv gpurun_out/ncu_${tag}_src.csv gpurun_out/ncu_${tag}_cuda.csv
