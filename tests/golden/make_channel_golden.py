"""Generate golden f-vectors from the REFERENCE ChannelSampler (run in the build container only).

Loads ``/root/reference/src/tsim/noise/channels.py`` by file path (it only needs NumPy),
samples a few configurations and stores inputs + outputs so the product's own sampler
(``tsim_b200.noise.ChannelSampler``) can be checked without the reference present.

    python tests/golden/make_channel_golden.py
"""

import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/src/tsim/noise/channels.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref():
    spec = importlib.util.spec_from_file_location("_ref_channels", REF)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def dump(name, ref_sampler, calls, extra):
    sd = ref_sampler._sparse_data
    out = dict(extra)
    out["n_channels"] = np.array([len(sd)])
    out["num_f"] = np.array([ref_sampler.signature_matrix.shape[1]])
    for i, (p, cdf, pats) in enumerate(sd):
        out[f"p{i}"] = np.array([p])
        out[f"cdf{i}"] = cdf
        out[f"pat{i}"] = pats
    out["calls"] = np.array(calls)
    for j, n in enumerate(calls):
        out[f"f{j}"] = np.packbits(ref_sampler.sample(n), axis=1, bitorder="little")
    np.savez_compressed(os.path.join(HERE, name), **out)


def main():
    ch = load_ref()
    # 1. independent single-bit channels, identity transform (the bench's noise model)
    q = 1e-3 * (1 + (np.arange(63) % 15))
    s = ch.ChannelSampler([ch.error_probs(x) for x in q], np.eye(63, dtype=np.uint8), seed=12345)
    dump("channel_sampler_bits.npz", s, [1000, 257, 1], {"q": q, "seed": np.array([12345])})
    # 2. multi-bit channels through a dense transform (depolarising-like), exercises cond_cdf
    rng = np.random.default_rng(7)
    probs = [ch.pauli_channel_1_probs(0.01, 0.02, 0.03), ch.error_probs(0.05), ch.pauli_channel_2_probs(*([0.002] * 15))]
    n_e = sum(int(np.log2(len(p))) for p in probs)
    T = rng.integers(0, 2, size=(9, n_e)).astype(np.uint8)
    s = ch.ChannelSampler(probs, T, seed=99)
    dump("channel_sampler_pauli.npz", s, [500, 33], {"seed": np.array([99])})


if __name__ == "__main__":
    main()
