#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout -s KILL 600 python -m pytest tests/test_gpu_program_npz.py tests/test_gpu_noise.py -x -q 2>&1 | tail -4
timeout -s KILL 300 python tools/batch_sweep.py 2>&1 | tail -10
