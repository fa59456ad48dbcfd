#!/usr/bin/env python
"""Device-resident timing of the sliced path with and without the pattern cache (one workload, several weights)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2_distill35")
    ap.add_argument("--shots", type=int, default=1_000_000)
    ap.add_argument("--p", type=float, default=1e-3)
    ap.add_argument("--mode", default="sliced")
    ap.add_argument("--weights", default="off,0,1,2,3")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch

    from tsim_b200.backend import DeviceProgram
    from tsim_b200.noise import ChannelSampler
    from tsim_b200.synthetic import noise_probs, synthetic_program

    prog = synthetic_program(args.workload)
    cs = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), args.p), seed=12345)
    dp = DeviceProgram(prog, mode=args.mode, pattern_cache=None)
    wo = dp.info["words_out64"]
    B = args.shots
    fs = [torch.from_numpy(cs.sample_packed(B).view(np.int64)).cuda() for _ in range(4)]
    outs = [torch.empty((B, wo), dtype=torch.int64, device="cuda") for _ in range(4)]
    st = torch.cuda.current_stream().cuda_stream
    base = None
    for wtxt in args.weights.split(","):
        w = None if wtxt == "off" else int(wtxt)
        n = dp.set_pattern_cache(w)
        for i in range(3):
            dp.sample_device(fs[i % 4].data_ptr(), B, (1, i), outs[i % 4].data_ptr(), stream=st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.reps):
            dp.sample_device(fs[i % 4].data_ptr(), B, (1, i), outs[i % 4].data_ptr(), stream=st)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.reps
        dp.sample_device(fs[0].data_ptr(), B, (7, 7), outs[0].data_ptr(), stream=st)
        torch.cuda.synchronize()
        got = outs[0].cpu().numpy()
        if base is None:
            base = got
        same = bool(np.array_equal(got, base))
        print(f"{args.workload} {args.mode} cache={wtxt:>3} entries={n:>9} {ms:8.4f} ms/step {B / ms * 1e3:10.3e} shots/s same_bits={same}", flush=True)


if __name__ == "__main__":
    main()
