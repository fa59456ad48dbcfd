"""Host error-mechanism sampler vs golden f-vectors produced by the reference ChannelSampler."""

import os

import numpy as np
import pytest

from tsim_b200.noise import ChannelSampler, pack_f_rows

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tables(z):
    return [(float(z[f"p{i}"][0]), z[f"cdf{i}"], z[f"pat{i}"]) for i in range(int(z["n_channels"][0]))]


@pytest.mark.parametrize("name", ["channel_sampler_bits.npz", "channel_sampler_pauli.npz"])
def test_stream_matches_reference(name):
    z = np.load(os.path.join(GOLD, name))
    num_f = int(z["num_f"][0])
    s = ChannelSampler.from_sparse(_tables(z), num_f, seed=int(z["seed"][0]))
    for j, n in enumerate(z["calls"]):
        want = np.unpackbits(z[f"f{j}"], axis=1, bitorder="little", count=num_f)
        got = s.sample(int(n))
        assert got.dtype == np.uint8 and got.shape == (n, num_f)
        assert np.array_equal(got, want)


def test_from_bit_probs_equals_reference_construction():
    z = np.load(os.path.join(GOLD, "channel_sampler_bits.npz"))
    s = ChannelSampler.from_bit_probs(z["q"], seed=int(z["seed"][0]))
    want = np.unpackbits(z["f0"], axis=1, bitorder="little", count=63)
    assert np.array_equal(s.sample(int(z["calls"][0])), want)


@pytest.mark.parametrize("num_f", [1, 63, 64, 65, 130])
def test_packed_equals_dense(num_f):
    q = np.linspace(0.01, 0.2, num_f)
    a = ChannelSampler.from_bit_probs(q, seed=5)
    b = ChannelSampler.from_bit_probs(q, seed=5)
    for n in (0, 1, 77):
        dense = a.sample(n)
        packed = b.sample_packed(n)
        assert packed.shape == (n, max(1, (num_f + 63) // 64))
        assert np.array_equal(pack_f_rows(dense), packed)


def test_x_error_direct_detector_kat():
    # reference test/integration/test_sampler_circuits.py:25-37: X_ERROR(0.3), seed 1, 10 shots -> 4 ones.
    # Seed schedule of sampler.py:203: channel_seed = default_rng(seed).integers(0, 2**30)
    channel_seed = int(np.random.default_rng(1).integers(0, 2**30))
    s = ChannelSampler.from_bit_probs([0.3], seed=channel_seed)
    f = s.sample(10)
    assert int(f.sum()) == 4
    assert list(np.flatnonzero(f[:, 0])) == [1, 6, 8, 9]
