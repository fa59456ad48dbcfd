#!/usr/bin/env python
"""End-to-end time of tsim_b200.sampler.sample_program(uint8[B, num_f] host rows) for the input-path variants:
host-side packing (TSIM_B200_HOST_PACK=1, thread count) vs byte DMA + device packing, pinned vs pageable input."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch

    import tsim_b200.sampler as S
    from tsim_b200.backend import DeviceProgram, PinnedArray, split_key
    from tsim_b200.noise import ChannelSampler
    from tsim_b200.synthetic import noise_probs, synthetic_program

    name, shots = os.environ.get("WL", "cfg2_distill35"), int(os.environ.get("SHOTS", "1000000"))
    prog = synthetic_program(name)
    cs = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), 1e-3), seed=1)
    S.check_norm_deviations = lambda d: None
    dp = DeviceProgram(prog, pattern_cache=None if os.environ.get("MEMO", "0") == "0" else 3)
    f = cs.sample(shots)
    pin = PinnedArray(f.shape, np.uint8)
    pin.array[...] = f
    for label, arr in (("pinned", pin.array), ("pageable", f)):
        key = (1, 1)
        for _ in range(3):
            key, sub = split_key(key)
            S.sample_program(dp, arr, sub)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            key, sub = split_key(key)
            t0 = time.perf_counter()
            S.sample_program(dp, arr, sub)
            ts.append(time.perf_counter() - t0)
        print(f"  {label:8s} median {np.median(ts) * 1e3:7.3f} ms  min {min(ts) * 1e3:7.3f} ms  -> {shots / np.median(ts):.3e} shots/s", flush=True)


if __name__ == "__main__":
    if os.environ.get("E2E_CHILD"):
        child()
    else:
        for hp, th in (("0", "1"), ("1", "2"), ("1", "4"), ("1", "8"), ("1", "16")):
            for memo in ("0", "1"):
                env = dict(os.environ, E2E_CHILD="1", TSIM_B200_HOST_PACK=hp, TSIM_B200_HOST_THREADS=th, MEMO=memo)
                print(f"host_pack={hp} threads={th} memo={memo}", flush=True)
                subprocess.run([sys.executable, __file__], env=env)
