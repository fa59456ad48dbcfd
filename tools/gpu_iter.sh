#!/bin/bash
# One GPU-box iteration on the sliced kernel: sliced parity tests, bench (sliced), ncu full capture.  Logs in gpurun_out/.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -k sliced > gpurun_out/t_sliced.log 2>&1
echo "sliced tests rc=$?"; tail -3 gpurun_out/t_sliced.log
timeout -s KILL 600 python bench.py --no-cpu --no-extras --steps 10 --warmup 3 > gpurun_out/bench_sliced.json 2> gpurun_out/bench_sliced.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_sliced.json"))
    print("ms_per_step",d["ms_per_step"],"value",d["value"],"e2e",d["e2e"]["value"],"kernel_ms",d["roofline"]["kernel_ms"])
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/bench_sliced.err").read()[-2000:])
PY
if [ "$1" == "prof" ]; then
ncu --set full --clock-control none --import-source on -k regex:sample_sliced_kernel -s 3 -c 1 -o gpurun_out/sliced_full -f \
  python bench.py --mode sliced --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/prof_full.log 2>&1
fi
