#!/bin/bash
mkdir -p gpurun_out
for t in memcheck racecheck synccheck; do
  SANITIZER_QUICK=1 timeout -s KILL 1200 compute-sanitizer --tool $t --error-exitcode 9 python tools/sanitizer_smoke.py > gpurun_out/sanitizer_$t.log 2>&1; echo "$t rc=$?"; grep -E "SUMMARY|done" gpurun_out/sanitizer_$t.log | tail -2
done
