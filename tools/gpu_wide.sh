#!/bin/bash
# One GPU-box iteration on the wide sliced layout: its tests, the sliced parity tests, then narrow vs wide bench lines.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_sliced_wide.py -x -q > gpurun_out/t_wide.log 2>&1; echo "wide tests rc=$?"; tail -5 gpurun_out/t_wide.log
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -k "sliced and not wide" > gpurun_out/t_sliced.log 2>&1; echo "sliced tests rc=$?"; tail -3 gpurun_out/t_sliced.log
Q="--no-cpu --no-extras --no-configs --no-parity --no-sustain"
for w in 0 1; do
  TSIM_B200_SLICED_WIDE=$w timeout -s KILL 600 python bench.py --steps 10 --warmup 3 $Q > gpurun_out/bench_wide$w.json 2> gpurun_out/bench_wide$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_wide$w.json"))
    print("wide=$w ms_per_step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"memo",d["memoised"]["ms_per_step"],"e2e",d["e2e"]["value"])
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/bench_wide$w.err").read()[-1500:])
PY
done
