#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_pattern_cache.py -x -q > gpurun_out/t_memo.log 2>&1; echo "cache tests rc=$?"; tail -15 gpurun_out/t_memo.log
timeout -s KILL 300 python tools/memo_bench.py 2>&1 | tail -8
timeout -s KILL 300 python tools/memo_bench.py --workload cfg5_distill85 --shots 100000 2>&1 | tail -8
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/t_par.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/t_par.log
