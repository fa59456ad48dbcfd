"""K5 (device channel sampler) and the fused noise+sample pipeline.  `-m gpu`.

Parity with the reference's ChannelSampler is statistical (different random stream); tolerances follow the
reference's own tests (test/unit/noise/test_channels.py:987-1048 use rtol 5-10 % at 1e5 samples) but are stated
in standard deviations here."""

import os

import numpy as np
import pytest

import oracle
from tsim_b200.noise import ChannelSampler, DeviceChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _unpack(packed, n):
    return np.unpackbits(packed.view(np.uint8), axis=1, bitorder="little", count=n)


def test_bit_channels_have_the_right_rates():
    q = noise_probs(63, 1e-3)
    B = 2_000_000
    f = _unpack(DeviceChannelSampler.from_bit_probs(q, seed=7).sample_packed(B), 63)
    freq = f.mean(axis=0)
    sigma = np.sqrt(q * (1 - q) / B)
    assert np.all(np.abs(freq - q) < 5 * sigma)
    # independence of two channels: joint rate = product
    both = (f[:, 13] & f[:, 14]).mean()
    assert abs(both - q[13] * q[14]) < 5 * np.sqrt(q[13] * q[14] / B)


def test_xor_of_two_channels_and_multi_outcome_channels():
    # reference test_simple_xor_two_errors: f0 = e0 ^ e1, P = 0.2*0.7 + 0.3*0.8
    data = [(0.2, np.array([1.0]), np.array([[1]])), (0.3, np.array([1.0]), np.array([[1]]))]
    f = DeviceChannelSampler(data, 1, seed=42).sample(400_000)
    assert abs(f[:, 0].mean() - 0.38) < 5 * np.sqrt(0.38 * 0.62 / 400_000)
    # multi-bit channels through a dense transform: compare with the host sampler (reference stream) statistically
    z = np.load(os.path.join(GOLD, "channel_sampler_pauli.npz"))
    tables = [(float(z[f"p{i}"][0]), z[f"cdf{i}"], z[f"pat{i}"]) for i in range(int(z["n_channels"][0]))]
    nf = int(z["num_f"][0])
    B = 1_000_000
    dev = DeviceChannelSampler(tables, nf, seed=3).sample(B).astype(np.int64)
    host = ChannelSampler.from_sparse(tables, nf, seed=3).sample(B).astype(np.int64)
    for a, b in ((dev.mean(0), host.mean(0)), ((dev[:, :-1] & dev[:, 1:]).mean(0), (host[:, :-1] & host[:, 1:]).mean(0))):
        sig = np.sqrt(np.maximum(b, 1e-7) * 2 / B)
        assert np.all(np.abs(a - b) < 6 * sig)


def test_rows_are_a_function_of_seed_call_and_shot_index():
    q = noise_probs(130, 5e-3)  # three words per row
    s = DeviceChannelSampler.from_bit_probs(q, seed=11)
    full = s.sample_packed(5000, call=4)
    assert full.shape == (5000, 3)
    parts = [s.sample_packed(hi - lo, shot_offset=lo, call=4) for lo, hi in ((0, 1), (1, 2049), (2049, 5000))]
    assert np.array_equal(np.concatenate(parts), full)
    assert np.array_equal(DeviceChannelSampler.from_bit_probs(q, seed=11).sample_packed(5000, call=4), full)
    assert not np.array_equal(s.sample_packed(5000, call=5), full)
    assert not np.array_equal(DeviceChannelSampler.from_bit_probs(q, seed=12).sample_packed(5000, call=4), full)
    skipped = s.sample_packed(5000, call=4, skip_shot0=True)
    assert not skipped[0].any() and np.array_equal(skipped[1:], full[1:])
    assert s.sample_packed(0).shape == (0, 3)
    # no bits beyond num_f
    assert not (full[:, 2] >> np.uint64(130 - 128)).any()


@pytest.mark.parametrize("mode", ["fast", "faithful", "sliced"])
def test_fused_pipeline_is_bit_exact_given_its_own_f_rows(mode):
    from tsim_b200.backend import DeviceProgram

    prog = synthetic_program("cfg2_distill35")
    dp = DeviceProgram(prog, mode=mode)
    noise = DeviceChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), 2e-3), seed=5)
    B = 200_000  # two pipeline slices
    bits, dev, f = dp.sample_noisy(noise, B, (0, 9), return_f=True)
    fb = _unpack(f, prog.infer_num_f())
    assert 0.001 < fb.mean() < 0.05
    for lo in (0, 131_072 - 512, B - 1024):
        want, want_dev = oracle.sample_program(prog, fb[lo : lo + 1024], (0, 9), shot_offset=lo, return_deviations=True, check_norm=False)
        assert np.array_equal(bits[lo : lo + 1024], want)
        if lo == 0:
            assert np.array_equal(np.asarray(dev, np.float32), np.asarray(want_dev, np.float32))
    packed, _ = dp.sample_noisy(noise, B, (0, 9), call=0, packed_out=True)
    assert np.array_equal(_unpack(packed, prog.num_outputs).astype(bool), bits)


def test_detector_sampler_with_device_noise_matches_host_noise_statistically(monkeypatch):
    import tsim_b200.sampler as S

    monkeypatch.setattr(S, "check_norm_deviations", lambda devs: None)
    prog = synthetic_program("cfg2_distill35")
    q = noise_probs(prog.infer_num_f(), 1e-3)
    shots = 400_000
    host = S.CompiledDetectorSampler(prog, ChannelSampler.from_bit_probs(q, seed=1), seed=3)
    dev = S.CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(q, seed=1), seed=3)
    a = host.sample(shots, batch_size=200_000, append_observables=True)
    b = dev.sample(shots, batch_size=200_000, append_observables=True)
    assert a.shape == b.shape == (shots, prog.num_outputs) and b.dtype == np.bool_
    pa, pb = a.mean(0), b.mean(0)
    sig = np.sqrt(np.maximum(pa * (1 - pa), 1e-6) * 2 / shots)
    assert np.all(np.abs(pa - pb) < 6 * sig)
    # reference-sample plumbing on the fused path: row 0 of the first batch is noiseless and removed
    c = dev.sample(1000, batch_size=500, append_observables=True, use_detector_reference_sample=True)
    assert c.shape == (1000, prog.num_outputs)
