"""Print the per-slice timeline of one reference-facing sample_program call (TSIM_B200_TRACE=1)."""
import os, sys, time
os.environ["TSIM_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tsim_b200 import sampler as S
from tsim_b200.backend import DeviceProgram, PinnedArray
from tsim_b200.noise import ChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program

prog = synthetic_program("cfg2_distill35")
dp = DeviceProgram(prog)
B = 1_000_000
f = PinnedArray((B, dp.num_f), np.uint8)
f.array[...] = ChannelSampler.from_bit_probs(noise_probs(dp.num_f), seed=1).sample(B)
S.check_norm_deviations = lambda devs: None
for i in range(3):
    t0 = time.perf_counter()
    S.sample_program(dp, f.array, (0, i))
    print(f"call {i}: {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
