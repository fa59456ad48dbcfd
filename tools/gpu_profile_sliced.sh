#!/bin/bash
# ncu launch list + one full capture of the sliced sampling kernel (bench numbers printed under ncu are not bench values)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_sliced.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/prof_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sample_sliced -s 3 -c 1 -o gpurun_out/sliced_full -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/prof_full.log 2>&1
ls -la gpurun_out
