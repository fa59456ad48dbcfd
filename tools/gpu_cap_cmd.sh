#!/bin/bash
# ncu --set full capture of one launch of a kernel inside an arbitrary command; exports the raw + source CSV pages.
# usage: gpu_cap_cmd.sh NAME KERNEL_REGEX SKIP command...
mkdir -p gpurun_out
name=$1; rx=$2; skip=$3; shift 3
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/$name -f "$@" > gpurun_out/prof_$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
rm -f gpurun_out/$name.ncu-rep
ls -la gpurun_out | grep $name
