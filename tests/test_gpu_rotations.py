"""Approximate branch on the GPU: rotation-gate programs through ``tsb_evaluate_host`` / ``CompiledStateProbs`` /
the sampling kernels vs the reference's closed forms and the oracle.  `-m gpu`."""

import itertools

import numpy as np
import pytest

import oracle
from rotation_programs import ROT_CASES, expected_probability, joint_program, sampling_program
from tsim_b200.noise import ChannelSampler

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6
X8 = np.array(list(itertools.product([0, 1], repeat=3)), dtype=np.uint8)


@pytest.mark.parametrize("kind,angles", ROT_CASES)
def test_evaluate_host_matches_closed_form_and_oracle(kind, angles):
    from tsim_b200.backend import DeviceProgram

    prog = joint_program(kind, angles)
    dp = DeviceProgram(prog, joint=True)
    amp = dp.evaluate(0, 1, X8)
    want_amp = oracle.evaluate(prog.components[0].compiled_scalar_graphs[1], X8)
    assert np.array_equal(amp.view(np.uint32), want_amp.view(np.uint32))  # bit-identical complex64
    p = np.abs(amp) / np.abs(dp.evaluate(0, 0, X8[:, :1]))
    want = np.array([expected_probability(kind, angles, *x) for x in X8])
    big = want > 1e-9
    assert np.max(np.abs(p[big] - want[big]) / want[big]) <= REL_TOL
    assert np.max(np.abs(p[~big] - want[~big]), initial=0.0) <= 1e-7


@pytest.mark.parametrize("kind,angles", ROT_CASES)
def test_state_probs_matrix(kind, angles):
    """The reference's get_matrix check (test/helpers/util.py:19-25): 2 P(m0, m1) == |U|^2."""
    from tsim_b200.sampler import CompiledStateProbs

    sp = CompiledStateProbs(joint_program(kind, angles), ChannelSampler.from_bit_probs([0.0], seed=0), seed=0)
    mat = np.array([[2 * sp.probability_of(np.array([m0, m1]), batch_size=1)[0] for m1 in (0, 1)] for m0 in (0, 1)])
    want = np.array([[2 * expected_probability(kind, angles, 0, m0, m1) for m1 in (0, 1)] for m0 in (0, 1)])
    assert np.allclose(mat, want, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("mode", ["auto", "fast", "faithful", "sliced"])
@pytest.mark.parametrize("kind,angles", ROT_CASES[:5])
def test_rotation_sampling_bits_match_oracle(kind, angles, mode):
    from tsim_b200.backend import DeviceProgram

    prog = sampling_program(kind, angles)
    B = 100_000
    f = (np.random.default_rng(5).random((B, 1)) < 0.25).astype(np.uint8)
    dp = DeviceProgram(prog, mode=mode)
    got, dev = dp.sample(f, (0, 11))
    from oracle import cport

    want, want_dev = cport.sample_program(prog, f, (0, 11), return_deviations=True, threads=cport.max_threads())
    assert np.array_equal(got, want)
    assert np.array_equal(np.asarray(dev, np.float32), np.asarray(want_dev, np.float32)) and dev[0] < 1e-5
    for fv in (0, 1):
        sel = f[:, 0] == fv
        n = int(sel.sum())
        for m0, m1 in itertools.product((0, 1), repeat=2):
            p = expected_probability(kind, angles, fv, m0, m1)
            freq = np.count_nonzero(sel & (got[:, 0] == m0) & (got[:, 1] == m1)) / n
            assert abs(freq - p) <= 5 * np.sqrt(p * (1 - p) / n) + 1e-4
