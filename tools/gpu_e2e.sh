#!/bin/bash
mkdir -p gpurun_out
nproc
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -k "host_packed" 2>&1 | tail -3
timeout -s KILL 900 python tools/e2e_sweep.py 2>&1 | tee gpurun_out/e2e_sweep.txt | tail -40
