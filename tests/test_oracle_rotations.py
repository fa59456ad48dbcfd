"""Approximate branch (``evaluate.py:56-59``) pinned on the CPU: hand-built rotation-gate programs vs the closed forms the
reference's own tests assert (``test/integration/test_sampler_circuits.py:576-600, 603-635, 638-690``), and the margin
census of the float32 tail."""

import itertools

import numpy as np
import pytest

import oracle
from oracle import cport
from rotation_programs import ROT_CASES, expected_probability, joint_program, sampling_program

REL_TOL = 1e-6  # north_star: marginal probabilities within 1e-6 relative
X8 = np.array(list(itertools.product([0, 1], repeat=3)), dtype=np.uint8)  # all (f, m0, m1)


@pytest.mark.parametrize("kind,angles", ROT_CASES)
def test_rotation_marginals_match_closed_form(kind, angles):
    prog = joint_program(kind, angles)
    norm, joint = prog.components[0].compiled_scalar_graphs
    assert joint.prefactor.has_approximate_floatfactors  # the float32 branch is what runs
    p_norm = np.abs(oracle.evaluate(norm, X8[:, :1]))
    p = np.abs(oracle.evaluate(joint, X8)) / p_norm
    want = np.array([expected_probability(kind, angles, *x) for x in X8])
    big = want > 1e-9
    assert np.max(np.abs(p[big] - want[big]) / want[big]) <= REL_TOL
    assert np.max(np.abs(p[~big] - want[~big]), initial=0.0) <= 1e-7
    # rows of |U|^2 sum to one (unitarity), as the reference's get_matrix comparisons imply
    for f in (0, 1):
        for m0 in (0, 1):
            assert abs(sum(2 * p[4 * f + 2 * m0 + m1] for m1 in (0, 1)) - 1) <= 2e-6


@pytest.mark.parametrize("kind,angles", ROT_CASES[:5])
def test_rotation_sampling_frequencies(kind, angles):
    prog = sampling_program(kind, angles)
    B = 40_000
    f = (np.random.default_rng(5).random((B, 1)) < 0.25).astype(np.uint8)
    bits, devs = oracle.sample_program(prog, f, (0, 11), return_deviations=True)
    assert devs[0] < 1e-5  # a genuine probability tree: the reference would not warn
    assert np.array_equal(bits, cport.sample_program(prog, f, (0, 11)))
    for fv in (0, 1):
        sel = f[:, 0] == fv
        n = int(sel.sum())
        for m0, m1 in itertools.product((0, 1), repeat=2):
            want = expected_probability(kind, angles, fv, m0, m1)
            got = np.count_nonzero(sel & (bits[:, 0] == m0) & (bits[:, 1] == m1)) / n
            assert abs(got - want) <= 5 * np.sqrt(want * (1 - want) / n) + 1e-4


def test_margin_census_small():
    from tsim_b200.noise import ChannelSampler
    from tsim_b200.synthetic import noise_probs, synthetic_program

    prog = synthetic_program("cfg2_distill35")
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=12345).sample(20_000)
    c = cport.census(prog, f, (0, 42))
    assert c["draws"] == 20_000 * 5
    # every draw outside the margin band has the same bit under the float64 evaluation of the same amplitudes
    assert c["outside_margin_flips_f64"] == 0 and c["outside_wide_margin_flips_f64"] == 0
    assert c["margin_draws"] <= c["margin_draws_wide"] <= 100
    assert c["max_rel_dev_f32_vs_f64"] < 2.0**-10


@pytest.mark.parametrize("kind,angles", ROT_CASES[:5])
def test_margin_census_rotation_programs(kind, angles):
    prog = sampling_program(kind, angles)
    f = (np.random.default_rng(5).random((50_000, 1)) < 0.25).astype(np.uint8)
    c = cport.census(prog, f, (0, 11))
    assert c["outside_margin_flips_f64"] == 0
    assert c["max_rel_dev_f32_vs_f64"] < 1e-5
