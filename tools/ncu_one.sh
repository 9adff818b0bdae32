#!/bin/bash
# tools/ncu_one.sh <tag> <kernel-regex> <python args...>: one `ncu --set full` capture (last launch of 3), raw CSV -> gpurun_out/
tag=$1; shift; rex=$1; shift
timeout 300 ncu --set full --import-source on --clock-control none -k regex:$rex -s 2 -c 1 -f -o /tmp/ncu_$tag python "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/ncu_$tag.ncu-rep --page details --csv > gpurun_out/ncu_${tag}_details.csv 2>/dev/null
ls -la gpurun_out/ncu_${tag}_raw.csv
