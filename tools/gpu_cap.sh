#!/bin/bash
# ncu --set full capture of one sliced-kernel launch of the default bench; exports raw + source CSV pages (the report itself stays on the box).
# usage: gpu_cap.sh NAME [env assignments...]
mkdir -p gpurun_out
name=$1; shift
Q="--no-cpu --no-extras --no-configs --no-parity --no-sustain"
env "$@" ncu --set full --clock-control none --import-source on -k regex:sample_sliced -s 3 -c 1 -o gpurun_out/$name -f \
  python bench.py --steps 2 --warmup 3 $Q > gpurun_out/prof_$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
rm -f gpurun_out/$name.ncu-rep
ls -la gpurun_out | tail -5
