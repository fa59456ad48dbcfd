#!/bin/bash
# Round-end style pass: whole GPU suite, smoke, default bench (with cpu baseline + extras), reference arm.
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout -s KILL 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json
