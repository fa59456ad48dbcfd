"""Sweep the host pipeline slice size (TSIM_B200_SLICE) and print end-to-end throughput of sample_program (byte rows in, bools out)."""
import json, os, subprocess, sys

for s in (131072, 200000, 262144, 333334, 500000, 1000000):
    env = dict(os.environ, TSIM_B200_SLICE=str(s))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3", "--no-cpu", "--no-extras", "--no-configs", "--no-sustain", "--no-parity"],
                         env=env, capture_output=True, text=True).stdout.strip().splitlines()
    d = json.loads(out[-1])
    e = d["e2e"]
    print(s, f"e2e pinned {e['value']:.3e} pageable {e['pageable']:.3e} memoised {e['memoised']:.3e} memoised+pageable {e['memoised_pageable']:.3e}", flush=True)
