#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_pattern_cache.py -x -q > gpurun_out/t_cfg4.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_cfg4.log
timeout -s KILL 300 python tools/memo_bench.py --workload cfg4_cultivation_d3 --shots 1000000 --mode sliced --weights off,2 --reps 3 2>&1 | tail -4
timeout -s KILL 300 python tools/memo_bench.py --workload cfg4_cultivation_d3 --shots 125000 --mode sliced --weights off,2 --reps 3 2>&1 | tail -4
timeout -s KILL 300 python tools/memo_bench.py --workload cfg4_cultivation_d3 --shots 1000000 --mode fast --weights off --reps 3 2>&1 | tail -4
