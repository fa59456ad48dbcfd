#!/bin/bash
# stage-size sweep of the sliced ring (narrow layout), then the same for wide at the best two
mkdir -p gpurun_out
Q="--no-cpu --no-extras --no-configs --no-parity --no-sustain"
for cw in "$@"; do
  TSIM_B200_SLICED_CHUNK_WORDS=$cw timeout -s KILL 600 python bench.py --steps 10 --warmup 3 $Q > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json"))
    print("chunk_words=$cw ms_per_step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"memo",d["memoised"]["ms_per_step"])
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
done
