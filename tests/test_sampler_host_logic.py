"""Host logic of the sampler classes on the CPU: the device is replaced by an oracle-backed fake, the way the
reference's own unit tests replace `sample_program` / spy on `ChannelSampler.sample`
(test/unit/test_sampler.py:247-346, test/unit/test_postselection.py:192-283)."""

from unittest.mock import patch

import numpy as np
import pytest

import kat_programs as K
import oracle
import tsim_b200.sampler as S
from oracle import evaluation as E
from tsim_b200.noise import ChannelSampler
from tsim_b200.program import CompiledComponent, make_program
from tsim_b200.shard import pack_bool_rows
from tsim_b200.synthetic import random_level


class FakeDeviceProgram:
    """Same surface as tsim_b200.backend.DeviceProgram, computed by the oracle."""

    calls = []

    def __init__(self, program, *, device=0, mode="auto", joint=False):
        self.program, self.joint, self.device = program, joint, device
        n_comp = len(program.components)
        self.info = {"n_components": n_comp, "words_out64": max(1, (program.num_outputs + 63) // 64), "data_bytes": 0,
                     "resident": 1, "mode": 1, "num_f": program.infer_num_f(), "num_outputs": program.num_outputs}
        self.num_outputs = program.num_outputs
        self.num_f = program.infer_num_f()

    def sample(self, f, key, *, shot_offset=0, packed_out=False, out=None):
        f = np.asarray(f)
        if f.dtype == np.uint64:
            f = np.unpackbits(f.view(np.uint8), axis=1, bitorder="little", count=max(self.num_f, 0))[:, : self.num_f]
        FakeDeviceProgram.calls.append((f.shape[0], tuple(key)))
        bits, devs = oracle.sample_program(self.program, f, key, shot_offset=shot_offset, return_deviations=True, check_norm=False)
        return (pack_bool_rows(bits) if packed_out else bits), np.asarray(devs, np.float32)

    def evaluate(self, component, level, params):
        return E.evaluate(self.program.components[component].compiled_scalar_graphs[level], np.asarray(params))


@pytest.fixture(autouse=True)
def fake_device(monkeypatch):
    FakeDeviceProgram.calls = []
    monkeypatch.setattr(S, "DeviceProgram", FakeDeviceProgram)
    # sample_program's handle cache would look for a real DeviceProgram: route the seam to the fake as well
    monkeypatch.setattr(S, "device_program_for", lambda p, **kw: p if isinstance(p, FakeDeviceProgram) else FakeDeviceProgram(p))


def _hm(seed=0):
    return S.CompiledMeasurementSampler(K.hm_program(), ChannelSampler.from_bit_probs([], seed=0), seed=seed)


def test_seed_chain_through_the_sampler_class():
    # reference test/unit/test_sampler.py:223-233, now through batching + key schedule of the class
    s = _hm(0)
    assert [int(np.count_nonzero(s.sample(100))) for _ in range(4)] == [48, 53, 52, 50]


@pytest.mark.parametrize(("shots", "expected_batch_size"), [(100, 25), (101, 26)])
def test_auto_batch(shots, expected_batch_size):
    # reference test/unit/test_sampler.py:247-274
    s = _hm(42)
    with patch.object(type(s), "_estimate_batch_size", return_value=30), patch.object(
        s._channel_sampler, "sample_packed", wraps=s._channel_sampler.sample_packed
    ) as spy:
        out = s.sample(shots)
    assert out.shape == (shots, 1)
    assert [c.args[0] for c in spy.call_args_list] == [expected_batch_size] * 4


@pytest.mark.parametrize(
    ("shots", "max_batch", "batch_size", "compute_ref", "expected_bs", "expected_n"),
    [
        (99, 30, None, True, 25, 4),
        (100, 30, None, True, 26, 4),
        (200, 30, None, True, 29, 7),
        (100, 30, None, False, 25, 4),
        (100, None, 50, True, 51, 2),
        (100, None, 51, True, 51, 2),
        (2, None, 1, True, 2, 2),
        (12, None, 3, True, 4, 4),
    ],
)
def test_batch_size_with_reference(shots, max_batch, batch_size, compute_ref, expected_bs, expected_n):
    # reference test/unit/test_sampler.py:288-346 (same table)
    s = _hm(0)
    with patch.object(type(s), "_estimate_batch_size", return_value=max_batch or 9999), patch.object(
        s._channel_sampler, "sample_packed", wraps=s._channel_sampler.sample_packed
    ) as spy:
        result = s._sample_batches(shots, batch_size=batch_size, compute_reference=compute_ref)
    if compute_ref:
        samples, ref = result
        assert samples.shape == (shots, 1) and ref.shape == (1,)
    else:
        assert result.shape == (shots, 1)
    sizes = [c.args[0] for c in spy.call_args_list]
    assert sizes == [expected_bs] * expected_n
    assert [n for n, _ in FakeDeviceProgram.calls] == [expected_bs] * expected_n


def _det_program():
    """``R 0 1 2; X 2; M 0 1 2; DETECTOR rec[-2]; DETECTOR rec[-3]; OBSERVABLE_INCLUDE(0) rec[-1]`` plus one coin
    component, so that the device path is exercised: outputs [det0, det1, obs0(=1, flipped direct), coin]."""
    return make_program(
        [K.coin_component(3)],
        direct_f_indices=[0, 1, 2],
        direct_flips=[False, False, True],
        output_order=[0, 1, 2, 3],
        num_outputs=4,
        num_detectors=2,
        num_f=3,
    )


def _det_sampler(seed=0, probs=(0.0, 0.0, 0.0)):
    return S.CompiledDetectorSampler(_det_program(), ChannelSampler.from_bit_probs(list(probs), seed=1), seed=seed)


def test_reference_sample_flags():
    # reference test/unit/test_sampler.py:349-452
    full = _det_sampler().sample(5, append_observables=True)
    assert np.array_equal(full[:, :3], np.array([[0, 0, 1]] * 5))
    d, o = _det_sampler().sample(5, separate_observables=True, use_detector_reference_sample=True)
    assert not d.any() and np.array_equal(o[:, 0], np.ones(5, bool))
    d2, o2 = _det_sampler().sample(5, separate_observables=True, use_observable_reference_sample=True)
    assert not d2.any() and not o2[:, 0].any()  # the deterministic observable bit is XORed away
    packed = _det_sampler().sample(1, append_observables=True, bit_packed=True, use_detector_reference_sample=True,
                                   use_observable_reference_sample=True)
    plain = _det_sampler().sample(1, append_observables=True, use_detector_reference_sample=True,
                                  use_observable_reference_sample=True)
    assert np.array_equal(packed, np.packbits(plain, axis=1, bitorder="little"))
    a = _det_sampler().sample(7, append_observables=True)
    b = _det_sampler().sample(7, append_observables=True, use_detector_reference_sample=False, use_observable_reference_sample=False)
    assert np.array_equal(a, b)


def test_shapes_errors_and_layouts():
    s = _det_sampler()
    assert s.sample(0).shape == (0, 2)
    assert s.sample(0, append_observables=True).shape == (0, 4)
    d, o = s.sample(0, separate_observables=True)
    assert d.shape == (0, 2) and o.shape == (0, 2)
    assert s.sample(9, bit_packed=True).shape == (9, 1)
    with pytest.raises(ValueError, match="shots must be non-negative"):
        s.sample(-1)
    with pytest.raises(ValueError, match="batch_size must be at least 1"):
        s.sample(3, batch_size=0)
    with pytest.raises(ValueError, match="separate_observables"):
        s.sample(3, separate_observables=True, prepend_observables=True)
    with pytest.raises(ValueError, match="postselection_mask must have shape"):
        s.sample(3, postselection_mask=np.zeros(5, bool))
    base = _det_sampler(3).sample(20, append_observables=True)
    both = _det_sampler(3).sample(20, append_observables=True, prepend_observables=True)
    assert np.array_equal(both, np.concatenate([base[:, 2:], base[:, :2], base[:, 2:]], axis=1))
    assert "CompiledDetectorSampler(3 direct, 2 graphs" in repr(s)


def test_postselection_skips_the_device_for_discarded_shots():
    # reference test/unit/test_postselection.py:192-283: only survivors reach sample_program, in fixed-size batches
    mask = np.array([True, False])
    s = _det_sampler(5, probs=(0.5, 0.1, 0.0))
    seen = []
    real = S.sample_program

    def spy(program, f, key):
        seen.append(np.asarray(f).shape[0])
        return real(program, f, key)

    with patch.object(S, "sample_program", side_effect=spy):
        out = s.sample(200, batch_size=16, append_observables=True, postselection_mask=mask)
    discarded = out[:, 0]
    assert 40 < discarded.sum() < 160
    assert all(n == 16 for n in seen)  # fixed batch shape, the last batch is padded
    assert sum(seen) >= int((~discarded).sum()) and sum(seen) - 16 < int((~discarded).sum())
    assert not out[discarded][:, 3].any()  # component column stays False for discarded shots
    assert out[~discarded][:, 3].any()
    # discarded rows keep only their direct *detector* columns (sampler.py:524-526); the direct observable is left False
    assert out[~discarded][:, 2].all() and not out[discarded][:, 2].any()
    # a mask that touches no direct detector falls back to plain batching
    s2 = _det_sampler(5, probs=(0.5, 0.1, 0.0))
    plain = s2.sample(50, batch_size=16, postselection_mask=np.array([False, False]))
    assert plain.shape == (50, 2)


def test_norm_deviation_thresholds(monkeypatch):
    # sampler.py:149-161: ValueError when the deviation is ~1, UserWarning above 1e-5
    s = _hm(0)
    monkeypatch.setattr(FakeDeviceProgram, "sample", lambda self, f, key, **kw: (np.zeros((f.shape[0], 1), bool), np.array([1.0], np.float32)))
    with pytest.raises(ValueError, match="vanishing marginal"):
        s.sample(4)
    monkeypatch.setattr(FakeDeviceProgram, "sample", lambda self, f, key, **kw: (np.zeros((f.shape[0], 1), bool), np.array([3e-4], np.float32)))
    with pytest.warns(UserWarning, match="not normalized correctly"):
        s.sample(4)


def test_probability_of_matches_closed_form():
    rng = np.random.default_rng(4)
    F, n = 5, 2
    joint = random_level(rng, 3, F + n, 3, 2, 2, 1, approx=False, density=0.3)
    norm = random_level(rng, 2, F, 2, 1, 1, 0, approx=False, density=0.3)
    comp = CompiledComponent((0, 1), np.arange(F, dtype=np.int32), (norm, joint))
    prog = make_program([comp], num_outputs=2, num_f=F)
    sp = S.CompiledStateProbs(prog, ChannelSampler.from_bit_probs([0.2] * F, seed=3), seed=0)
    state = np.array([1, 0], dtype=np.uint8)
    got = sp.probability_of(state, batch_size=32)
    f = ChannelSampler.from_bit_probs([0.2] * F, seed=3).sample(32)
    with np.errstate(all="ignore"):
        want = E.evaluate_abs(joint, np.hstack([f, np.tile(state, (32, 1))])) / E.evaluate_abs(norm, f)
    np.testing.assert_allclose(got, want, rtol=1e-6, equal_nan=True)
    with pytest.raises(ValueError):
        sp.probability_of(state, batch_size=0)
    with pytest.raises(ValueError):
        sp.probability_of(np.zeros(3), batch_size=4)


def test_packed_column_slicing_matches_packbits():
    """`_packed_columns` (bit_packed=True path) == np.packbits of the bool columns, for any range and row width."""
    from tsim_b200 import sampler as S

    rng = np.random.default_rng(0)
    for W, n_out in ((1, 20), (1, 64), (2, 100), (3, 160), (2, 128)):
        bits = rng.integers(0, 2, size=(257, n_out)).astype(bool)
        pk = np.packbits(bits, axis=1, bitorder="little")
        tmp = np.zeros((257, 8 * W), np.uint8)
        tmp[:, : pk.shape[1]] = pk
        rows = tmp.view(np.uint64).copy()
        for _ in range(120):
            lo = int(rng.integers(0, n_out + 1))
            hi = int(rng.integers(lo, n_out + 1))
            want = np.packbits(bits[:, lo:hi], axis=1, bitorder="little")
            got = S._packed_columns(rows, lo, hi)
            assert got.dtype == np.uint8 and got.shape == want.shape and np.array_equal(got, want), (W, n_out, lo, hi)
