"""Post-selection with the survivor buffering on the device (tsb_postselect) vs the host buffering that mirrors the
reference line by line (src/tsim/sampler.py:422-545; expectations of test/unit/test_postselection.py:192-283).  `-m gpu`."""

import warnings

import numpy as np
import pytest

from test_gpu_parity import _random_program
from tsim_b200.noise import ChannelSampler, DeviceChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program

pytestmark = pytest.mark.gpu


def _sampler(prog, q, seed, *, noise_seed=6, device_noise=False, mode="auto"):
    from tsim_b200.sampler import CompiledDetectorSampler

    cs = (DeviceChannelSampler if device_noise else ChannelSampler).from_bit_probs(q, seed=noise_seed)
    return CompiledDetectorSampler(prog, cs, seed=seed, mode=mode)


def _mask_on_direct(prog, nd, k=2, return_f=False):
    """Mask on the first k direct detector columns whose bit is an unflipped f bit (a flipped one fires almost always)."""
    n_direct = len(prog.direct_f_indices)
    cols = [(int(c), int(fi)) for c, fi, fl in zip(np.asarray(prog.output_order[:n_direct]), prog.direct_f_indices, prog.direct_flips) if c < nd and not fl][:k]
    mask = np.zeros(nd, bool)
    mask[[c for c, _ in cols]] = True
    return (mask, [fi for _, fi in cols]) if return_f else mask


@pytest.mark.parametrize("mode", ["auto", "fast", "faithful"])
@pytest.mark.parametrize("shots,batch", [(200, 32), (1000, 64), (97, 97), (513, 1), (300, 1000)])
@pytest.mark.parametrize("ref_flags", [(False, False), (True, False), (True, True), (False, True)])
def test_device_buffering_equals_host_buffering(shots, batch, ref_flags, mode, monkeypatch):
    import tsim_b200.sampler as S

    monkeypatch.setattr(S, "check_norm_deviations", lambda devs: None)
    if batch == 1 and mode != "auto":
        pytest.skip("one batch-size-1 sweep is enough")
    prog = _random_program(9, n_comp=2, n_c=3, n_direct=4)
    prog.num_detectors = 6
    q = np.full(prog.infer_num_f(), 0.1)
    mask = _mask_on_direct(prog, 6)
    kw = dict(batch_size=batch, append_observables=True, postselection_mask=mask,
              use_detector_reference_sample=ref_flags[0], use_observable_reference_sample=ref_flags[1])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        monkeypatch.setenv("TSIM_B200_POSTSELECT", "host")
        host = _sampler(prog, q, 2, mode=mode)
        want = host.sample(shots, **kw)
        want2 = host.sample(shots, **kw)  # second call: key and noise streams carried on
        monkeypatch.setenv("TSIM_B200_POSTSELECT", "device")
        dev = _sampler(prog, q, 2, mode=mode)
        got = dev.sample(shots, **kw)
        got2 = dev.sample(shots, **kw)
    assert got.shape == want.shape == (shots, prog.num_outputs)
    assert np.array_equal(got, want), f"{np.count_nonzero((got != want).any(axis=1))} rows differ"
    assert np.array_equal(got2, want2)
    assert host._key == dev._key  # same number of dispatches (one key split each)


def test_session_counts_and_discards(monkeypatch):
    """Only survivors reach the sampling kernel, in batches of exactly batch_size (last one padded)."""
    import tsim_b200.sampler as S

    monkeypatch.setattr(S, "check_norm_deviations", lambda devs: None)
    prog = synthetic_program("cfg2_distill35")
    nd = prog.num_detectors
    q = noise_probs(prog.infer_num_f(), 2e-2)
    mask = _mask_on_direct(prog, nd, k=3)
    s = _sampler(prog, q, 5)
    sessions = []
    orig = s._device_program.postselect_session

    def spy(*a, **k):
        sessions.append(orig(*a, **k))
        return sessions[-1]

    monkeypatch.setattr(s._device_program, "postselect_session", spy)
    shots, batch = 50_000, 4096
    det, obs = s.sample(shots, batch_size=batch, separate_observables=True, postselection_mask=mask)
    discarded = (det & mask).any(axis=1)
    n_surv = int((~discarded).sum())
    assert 0 < n_surv < shots
    assert sessions[0].dispatches == -(-n_surv // batch)
    # discarded shots: direct detector columns only, everything else False (sampler.py:430-433)
    direct_cols = np.asarray(prog.output_order[: len(prog.direct_f_indices)])
    full = np.concatenate([det, obs], axis=1)
    other = np.setdiff1d(np.arange(prog.num_outputs), direct_cols[direct_cols < nd])
    assert not full[discarded][:, other].any()
    # kept shots have all masked detectors quiet and carry sampled (non-direct) columns
    assert not (det[~discarded] & mask).any()
    comp_cols = np.setdiff1d(np.arange(prog.num_outputs), direct_cols)
    assert full[~discarded][:, comp_cols].any()


def test_device_noise_postselection_runs_and_is_deterministic(monkeypatch):
    import tsim_b200.sampler as S

    monkeypatch.setattr(S, "check_norm_deviations", lambda devs: None)
    prog = synthetic_program("cfg2_distill35")
    nd = prog.num_detectors
    q = noise_probs(prog.infer_num_f(), 1e-2)
    mask, f_idx = _mask_on_direct(prog, nd, k=2, return_f=True)
    a = _sampler(prog, q, 3, device_noise=True).sample(30_000, batch_size=8192, postselection_mask=mask, append_observables=True)
    b = _sampler(prog, q, 3, device_noise=True).sample(30_000, batch_size=8192, postselection_mask=mask, append_observables=True)
    assert np.array_equal(a, b)
    rate = (a[:, :nd] & mask).any(axis=1).mean()
    want = 1 - np.prod(1 - np.asarray(q)[f_idx])
    assert abs(rate - want) < 5 * np.sqrt(want * (1 - want) / 30_000) + 1e-3
