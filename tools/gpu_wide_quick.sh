#!/bin/bash
# quick iteration: wide tests + one bench line (wide default)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_sliced_wide.py -x -q > gpurun_out/t_wide.log 2>&1; echo "wide tests rc=$?"; tail -3 gpurun_out/t_wide.log
Q="--no-cpu --no-extras --no-configs --no-parity --no-sustain"
for w in "$@"; do
  env $w timeout -s KILL 600 python bench.py --steps 10 --warmup 3 $Q > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json"))
    print("$w ms_per_step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"memo",d["memoised"]["ms_per_step"],"e2e",d["e2e"]["value"])
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
done
