#!/bin/bash
# One GPU-box pass for the sliced kernel: parity tests, then short benches.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/t_sliced.log 2>&1
echo "sliced tests rc=$?" | tee -a gpurun_out/t_sliced.log
tail -15 gpurun_out/t_sliced.log
for m in sliced fast; do
  timeout -s KILL 600 python bench.py --mode $m --no-cpu --no-extras --steps 10 --warmup 3 > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  echo "bench $m rc=$?"; cat gpurun_out/bench_$m.json | head -c 3000; tail -5 gpurun_out/bench_$m.err
done
