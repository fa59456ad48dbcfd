#!/bin/bash
# A/B of kernel variants: for each library given, run the sliced parity subset (first only) and the sliced bench.
mkdir -p gpurun_out
first=1
for lib in "$@"; do
  export TSIM_B200_LIB=$PWD/$lib
  if [ $first == 1 ]; then
    timeout -s KILL 900 python -m pytest tests -m gpu -x -q -k "sliced" > gpurun_out/t_sliced.log 2>&1
    echo "sliced tests ($lib) rc=$?"; tail -3 gpurun_out/t_sliced.log; first=0
  fi
  n=$(basename $lib .so)
  timeout -s KILL 600 python bench.py --mode sliced --no-cpu --no-extras --steps 10 --warmup 3 > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$n.json"))
    print("$n: ms_per_step %.4f value %.4g e2e %.4g kernel_ms %.4f" % (d["ms_per_step"],d["value"],d["e2e"]["value"],d["roofline"]["kernel_ms"]))
except Exception as e:
    print("$n: bench parse failed",e); print(open("gpurun_out/bench_$n.err").read()[-1500:])
PY
done
