#!/usr/bin/env python
"""K1s time against the structure of the program (no real compiled program can be produced
here, so the sensitivity of the pack-time algebra and of the kernel to structure is measured on synthetic variants of the
headline workload): mask density, masks shared by all graphs of a level, twice as many graphs.  Markdown table on stdout."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

VARIANTS = [
    ("base (density 0.15, i.i.d. masks)", dict()),
    ("density 0.05", dict(density=0.05)),
    ("density 0.30", dict(density=0.30)),
    ("masks shared by the graphs of a level", dict(shared_masks=True)),
    ("shared masks, density 0.30", dict(shared_masks=True, density=0.30)),
    ("graphs x 2", dict(graph_scale=2)),
]


def pack_stats(prog):
    """(parities, raw row loads, row loads the kernel issues incl. class padding) per slab of 32 shots, counted from the
    packer's term lists and from the records it emits."""
    from tsim_b200 import pack_sliced as ps

    terms_seen = []
    orig = ps._emit_runs

    def spy(terms, zero_row, scale=1, rotate=0, compact=True):
        terms_seen.append(list(terms))
        return orig(terms, zero_row, scale, rotate, compact)

    ps._emit_runs = spy
    pad = 0
    compact = ps.compact_items_pay([lv for c in prog.components for lv in c.compiled_scalar_graphs])
    if os.environ.get("TSIM_B200_SLICED_COMPACT"):
        compact = os.environ["TSIM_B200_SLICED_COMPACT"] != "0"
    try:
        for comp in prog.components:
            for lv in comp.compiled_scalar_graphs:
                recs, _, _ = ps.sliced_level_records(lv, 126, 127, compact=compact)
                for r, _t in recs:
                    o, end = ps.SLICED_HEADER_WORDS, ps.SLICED_HEADER_WORDS + (int(r[0]) & 0xFFFF)
                    while o < end:
                        kind, count = int(r[o]) & 0xFFFF, int(r[o]) >> 16
                        o += 4
                        if kind < 3 or 9 <= kind < 12:  # LIN / LIN2
                            nw = ps.CLASS_WORDS[kind % 3]
                            pad += 4 * nw * count
                            o += (4 if nw <= 3 else 8) * count
                        elif 3 <= kind < 9:  # PI
                            c1, c2 = ps.PI_CLASSES[kind - 3]
                            pad += 4 * (ps.CLASS_WORDS[c1] + ps.CLASS_WORDS[c2]) * count
                            o += 8 * count
                        elif 12 <= kind < 15:  # PAIR
                            pad += 8 * ps.CLASS_WORDS[kind - 12] * count
                            o += 12 * count
                        elif kind in (16, 17):
                            pad += 4 * count
                            o += 2 * count
                        elif 18 <= kind < 21:
                            pad += 4 * (1 + kind - 17) * count
                            o += 4 * count
                        elif kind == 21:
                            pad += 8 * count
                            o += 4 * count
                        else:  # generic block stream: count = words
                            e2 = o + count
                            while o < e2:
                                n = int(r[o + 1])
                                pad += 4 * n
                                o += 2 + n
    finally:
        ps._emit_runs = orig
    n_par = raw = 0
    for terms in terms_seen:
        for t in terms:
            for rows in ([t[2]] if t[0] == "lin" else [t[3], t[4]]):
                n_par += 1
                raw += len(rows)
    return n_par, raw, pad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2_distill35")
    ap.add_argument("--shots", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="", help="comma-separated variant indices (default: all)")
    args = ap.parse_args()
    import torch

    from tsim_b200.backend import DeviceProgram
    from tsim_b200.noise import ChannelSampler
    from tsim_b200.synthetic import noise_probs, synthetic_program

    B = args.shots
    print(f"workload {args.workload}, {B} shots per launch, device-resident packed rows, pattern cache off / on\n")
    print("| variant | parities | row loads (raw) | row loads (padded) | data KB | chunks | K1s ms | step ms | memoised step ms | K1s ns per padded load |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    base_ms = None
    pick = [int(v) for v in args.only.split(",") if v != ""]
    for vi, (label, kw) in enumerate(VARIANTS):
        if pick and vi not in pick:
            continue
        prog = synthetic_program(args.workload, **kw)
        n_par, raw, pad = pack_stats(prog)
        cs = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=12345)
        dp = DeviceProgram(prog, mode="sliced", pattern_cache=None)
        wo = dp.info["words_out64"]
        fs = [torch.from_numpy(cs.sample_packed(B).view(np.int64)).cuda() for _ in range(4)]
        outs = [torch.empty((B, wo), dtype=torch.int64, device="cuda") for _ in range(4)]
        st = torch.cuda.current_stream().cuda_stream

        def run(reps):
            for i in range(3):
                dp.sample_device(fs[i % 4].data_ptr(), B, (1, i), outs[i % 4].data_ptr(), stream=st)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(reps):
                dp.sample_device(fs[i % 4].data_ptr(), B, (1, i), outs[i % 4].data_ptr(), stream=st)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        step = run(args.reps)
        k_ms = []
        for i in range(5):
            dp.sample_device(fs[i % 4].data_ptr(), B, (2, i), outs[i % 4].data_ptr(), stream=st)
            torch.cuda.synchronize()
            k_ms.append(dp.last_kernel_ms()[0])
        k1s = float(np.mean(k_ms))
        dp.set_pattern_cache(3)
        memo = run(args.reps)
        slabs_per_launch = B / 32
        print(f"| {label} | {n_par} | {raw} | {pad} | {dp.info['data_bytes'] / 1024:.0f} | {dp.info['n_chunks']} | {k1s:.3f} | {step:.3f} | {memo:.3f} | "
              f"{k1s * 1e6 / (pad * slabs_per_launch / 1e3):.3f} |", flush=True)
        dp.close()


if __name__ == "__main__":
    main()
