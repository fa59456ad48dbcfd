#!/bin/bash
# round 2: launch list of a short default bench + full ncu captures of the dominant kernels (numbers printed under ncu are not bench values).
# The reports are exported to CSV pages on the box (raw metrics, per-source-line) and removed: gpurun_out/ must stay under 64 MiB.
mkdir -p gpurun_out
Q="--no-cpu --no-extras --no-configs --no-parity --no-sustain"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2f_launches.csv \
  python bench.py --steps 2 --warmup 3 $Q > gpurun_out/prof_list.log 2>&1
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/$name -f "$@" > gpurun_out/prof_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap r2f_k1s_cfg2 sample_sliced 3 python bench.py --steps 2 --warmup 3 $Q
cap r2f_k1s_cfg4 sample_sliced 2 python tools/memo_bench.py --workload cfg4_cultivation_d3 --shots 1000000 --mode sliced --weights off --reps 2
cap r2f_light_cfg2 light_kernel 2 python tools/memo_bench.py --workload cfg2_distill35 --shots 1000000 --mode sliced --weights 3 --reps 2
ls -la gpurun_out | tail -12
