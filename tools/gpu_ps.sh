#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_postselect.py tests/test_gpu_parity.py -x -q > gpurun_out/t_ps.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_ps.log
