#!/bin/bash
# Round-end evidence in one call: full GPU suite, smoke, default bench + reference arm, structure sweep, launch list and
# ncu captures of the dominant kernels (CSV pages only; numbers printed under ncu are not bench values).
mkdir -p gpurun_out
bash tools/gpu_full.sh
python tools/structure_sweep.py > gpurun_out/structure_sweep.md 2>gpurun_out/structure_sweep.err; tail -8 gpurun_out/structure_sweep.md
bash tools/gpu_profile_r2.sh > gpurun_out/profile.log 2>&1; tail -3 gpurun_out/profile.log
python tools/sample_api_bench.py > gpurun_out/sample_api_cfg2.txt 2>&1; python tools/sample_api_bench.py cfg3_surface_d5 10000000 > gpurun_out/sample_api_cfg3.txt 2>&1; tail -4 gpurun_out/sample_api_cfg2.txt gpurun_out/sample_api_cfg3.txt
python tools/e2e_trace.py 2> gpurun_out/e2e_trace.txt; tail -12 gpurun_out/e2e_trace.txt
