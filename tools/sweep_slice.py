"""Sweep the host pipeline slice size (TSIM_B200_SLICE) and print device / end-to-end throughput."""
import json, os, subprocess, sys

for s in (131072, 200000, 262144, 333334, 500000, 1000000):
    env = dict(os.environ, TSIM_B200_SLICE=str(s))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3", "--no-cpu", "--no-extras"],
                         env=env, capture_output=True, text=True).stdout.strip().splitlines()
    d = json.loads(out[-1])
    print(s, f"device {d['value']:.3e}", f"e2e {d['e2e']['value']:.3e}", flush=True)
