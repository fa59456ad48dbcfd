"""e2e time of sample_program (pinned uint8 f in, bool out) against the pipeline slice size, pattern cache off / on."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, time, numpy as np
sys.path.insert(0, %r)
from tsim_b200 import sampler as S
from tsim_b200.backend import DeviceProgram, PinnedArray
from tsim_b200.noise import ChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program
prog = synthetic_program("cfg2_distill35")
S.check_norm_deviations = lambda devs: None
B = 1_000_000
for memo in (None, "default"):
    dp = DeviceProgram(prog, pattern_cache=memo)
    f = PinnedArray((B, dp.num_f), np.uint8)
    f.array[...] = ChannelSampler.from_bit_probs(noise_probs(dp.num_f), seed=1).sample(B)
    ts = []
    for i in range(12):
        t0 = time.perf_counter(); S.sample_program(dp, f.array, (0, i)); ts.append(time.perf_counter() - t0)
    t = float(np.median(ts[3:]))
    print("slice", sys.argv[1], "memo", memo, "%%.3f ms  %%.3e shots/s" %% (1e3 * t, B / t), flush=True)
''' % ROOT
for sl in sys.argv[1:] or ["131072", "196608", "262144", "349526", "524288"]:
    env = dict(os.environ, TSIM_B200_SLICE=sl)
    subprocess.run([sys.executable, "-c", CODE, sl], env=env)
