"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py
"""
import os, sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import oracle
from tsim_b200.backend import DeviceProgram
from tsim_b200.noise import ChannelSampler, DeviceChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program

QUICK = bool(os.environ.get("SANITIZER_QUICK"))  # cfg2 / sliced only (the default path and its helper kernels)
CONFIGS = (("cfg2_distill35", 1100),) if QUICK else (("cfg2_distill35", 1100), ("cfg3p_rank1", 300), ("cfg5_distill85", 600))
for name, B in CONFIGS:
    prog = synthetic_program(name)
    nf = prog.infer_num_f()
    f = ChannelSampler.from_bit_probs(noise_probs(nf, 5e-3), seed=1).sample(B)
    want = oracle.sample_program(prog, f, (1, 2), check_norm=False)
    for mode in (("sliced",) if QUICK else ("faithful", "fast", "sliced")):
        for limit in (None, "150000"):
            if limit:
                os.environ["TSIM_B200_SMEM_LIMIT"] = limit
            else:
                os.environ.pop("TSIM_B200_SMEM_LIMIT", None)
            dp = DeviceProgram(prog, mode=mode)
            got, dev = dp.sample(f, (1, 2))
            assert np.array_equal(got, want), (name, mode, limit)
            if mode != "sliced":
                dp.set_pattern_cache(1)
                got2, _ = dp.sample(f, (1, 2))
                assert np.array_equal(got2, want)
            noise = DeviceChannelSampler.from_bit_probs(noise_probs(nf, 5e-3), seed=3)
            bits, _, fp = dp.sample_noisy(noise, B, (4, 4), return_f=True)
            amp = dp.evaluate(0, 1, np.zeros((5, dp.level_params(0, 1)), np.uint8))
            dp.close()
    print(name, "ok", flush=True)
print("sanitizer smoke done")
