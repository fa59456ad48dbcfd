"""Output layout on the device (tsb_sample_noisy_host_layout) against the host NumPy form of the same flags.

With a device channel sampler, ``CompiledDetectorSampler.sample`` applies the detector / observable split, prepend / append,
the reference-sample XOR and bit packing on the GPU (reference flag ladder: src/tsim/sampler.py:791-868).  The host form
(``DEVICE_LAYOUT = False``) runs the same kernels and key / noise schedule and does the layout with NumPy, so the two must
agree bit for bit for every flag combination, single- and multi-batch.  `-m gpu`."""

import itertools

import numpy as np
import pytest

from tsim_b200.noise import DeviceChannelSampler
from tsim_b200.sampler import CompiledDetectorSampler
from tsim_b200.synthetic import noise_probs, synthetic_program

pytestmark = pytest.mark.gpu


@pytest.fixture
def no_norm_check(monkeypatch):
    import tsim_b200.sampler as S

    monkeypatch.setattr(S, "check_norm_deviations", lambda devs: None)


def _pair(name="cfg2_distill35"):
    prog = synthetic_program(name)
    probs = noise_probs(prog.infer_num_f())
    mk = lambda: CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(probs, seed=77), seed=5)
    a, b = mk(), mk()
    b.DEVICE_LAYOUT = False
    return prog, a, b


FLAGS = [
    dict(),
    dict(append_observables=True),
    dict(prepend_observables=True),
    dict(prepend_observables=True, append_observables=True),
    dict(separate_observables=True),
]


@pytest.mark.parametrize("bit_packed", [False, True])
@pytest.mark.parametrize("flags", FLAGS)
@pytest.mark.parametrize("ref", [(False, False), (True, False), (False, True), (True, True)])
def test_layout_flags_match_host_form(no_norm_check, flags, bit_packed, ref):
    prog, dev, host = _pair()
    kw = dict(flags, bit_packed=bit_packed, use_detector_reference_sample=ref[0], use_observable_reference_sample=ref[1])
    for shots, batch in ((1, None), (3000, None), (5000, 1250), (5000, 2048)):  # exact and ragged batch divisions
        got = dev.sample(shots, batch_size=batch, **kw)
        want = host.sample(shots, batch_size=batch, **kw)
        if isinstance(want, tuple):
            assert isinstance(got, tuple) and len(got) == 2
            for g, w in zip(got, want):
                assert g.dtype == w.dtype and g.shape == w.shape and np.array_equal(g, w)
        else:
            assert got.dtype == want.dtype and got.shape == want.shape
            assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} differing entries ({kw}, shots={shots}, batch={batch})"


def test_layout_on_wide_rows_and_direct_only_programs(no_norm_check):
    # cfg5: 45 outputs; cfg3: 121 direct outputs (two 64-bit words per row, no compiled component)
    for name in ("cfg5_distill85", "cfg3_surface_d5"):
        prog, dev, host = _pair(name)
        for kw in (dict(bit_packed=True), dict(separate_observables=True), dict(append_observables=True, bit_packed=True, use_detector_reference_sample=True)):
            got, want = dev.sample(4100, **kw), host.sample(4100, **kw)
            if isinstance(want, tuple):
                assert all(np.array_equal(g, w) and g.shape == w.shape for g, w in zip(got, want))
            else:
                assert got.shape == want.shape and np.array_equal(got, want)


def test_layout_argument_checks():
    from tsim_b200.backend import DeviceProgram

    prog = synthetic_program("cfg2_distill35")
    dp = DeviceProgram(prog)
    nz = DeviceChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=1)
    with pytest.raises(ValueError):
        dp.layout_row_bytes([(0, 1)] * 5, bit_packed=True)
    with pytest.raises(ValueError):
        dp.sample_noisy_layout(nz, 64, (0, 1), [(10, 50)], bit_packed=True)  # range beyond num_outputs
    assert dp.layout_row_bytes([(0, 15), (15, 5)], bit_packed=True, split=1) == (2, 1)
    assert dp.layout_row_bytes([(0, 15), (15, 5)], bit_packed=True) == (3, 0)
    assert dp.layout_row_bytes([(0, 15), (15, 5)], bit_packed=False) == (20, 0)
