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

QUICK = bool(os.environ.get("SANITIZER_QUICK"))  # sliced records only: cfg2 (the default path and its helper kernels) and cfg5 (heavy masks: generic runs, three f words)
CONFIGS = (("cfg2_distill35", 1100), ("cfg5_distill85", 600)) if QUICK else (("cfg2_distill35", 1100), ("cfg3p_rank1", 300), ("cfg5_distill85", 600))
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
            if mode == "sliced":
                # full evaluation (no pattern cache) through every variant of the sliced kernel: 8-way split (thin launch),
                # 4-way split (the headline kernel), 4-way wide layout; fused and separate transpose / assemble kernels
                dp.set_pattern_cache(None)
                for env in ({}, {"TSIM_B200_SLICED_SPLIT": "4"}, {"TSIM_B200_SLICED_SPLIT": "4", "TSIM_B200_SLICED_WIDE": "1"},
                            {"TSIM_B200_SLICED_SPLIT": "4", "TSIM_B200_SLICED_FUSE": "0"}):
                    os.environ.update(env)
                    got3, _ = dp.sample(f, (1, 2))
                    for k in env:
                        os.environ.pop(k)
                    assert np.array_equal(got3, want), (name, env)
                dp.set_pattern_cache(3)
            noise = DeviceChannelSampler.from_bit_probs(noise_probs(nf, 5e-3), seed=3)
            bits, _, fp = dp.sample_noisy(noise, B, (4, 4), return_f=True)
            lay, lay2, _, row0 = dp.sample_noisy_layout(noise, B, (4, 5), [(0, 3), (3, dp.num_outputs - 3)], bit_packed=True, split=1,
                                                        ref_mask=np.full(dp.info["words_out64"], 0xFFFF, np.uint64), skip_shot0=True)
            assert lay.shape[0] == B - 1 and row0 is not None
            amp = dp.evaluate(0, 1, np.zeros((5, dp.level_params(0, 1)), np.uint8))
            dp.close()
    print(name, "ok", flush=True)
print("sanitizer smoke done")
