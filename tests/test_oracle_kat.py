"""Pin the CPU oracle against the reference's own known-answer tests.

Every expected value below is a literal from a reference test (file:line cited);
no reference code runs here.
"""

import numpy as np
import pytest

import kat_programs as K
import oracle
from oracle import evaluation as E
from oracle.exact_scalar import ExactScalar
from oracle.threefry import key_from_seed, split, threefry2x32, uniform_f32
from tsim_b200.program import HalfPiPhases, NodePhases, PhasePairs, PiProducts, empty_scalar_graphs


def _run_batches(prog, seed, shots, n_batches):
    """Key schedule of reference sampler.py:198,399: one split per batch."""
    key = key_from_seed(seed)
    out = []
    for _ in range(n_batches):
        key, sub = split(key)
        f = np.zeros((shots, prog.infer_num_f()), np.uint8)
        out.append(oracle.sample_program(prog, f, sub))
    return out


def test_threefry_random123_kat():
    # Random123 kat_vectors: threefry2x32 20 rounds, pi digits
    o0, o1 = threefry2x32(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3)
    assert (int(o0), int(o1)) == (0xC4923A9C, 0x483DF7A0)
    o0, o1 = threefry2x32(0, 0, 0, 0)
    assert (int(o0), int(o1)) == (0x6B200159, 0x99BA4EFE)
    o0, o1 = threefry2x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert (int(o0), int(o1)) == (0x1CB996FC, 0xBB002BE7)


def test_uniform_range():
    u = uniform_f32((0, 42), 10000)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.02


def test_seed_chain_hm():
    # reference test/unit/test_sampler.py:223-233
    counts = [int(np.count_nonzero(x)) for x in _run_batches(K.hm_program(), 0, 100, 4)]
    assert counts == [48, 53, 52, 50]


def test_bell_state():
    # reference test/integration/test_sampler_circuits.py:10-22
    m = _run_batches(K.bell_program(), 0, 100, 1)[0]
    assert np.array_equal(m[:, 0], m[:, 1])
    assert np.count_nonzero(m[:, 0]) == 48


def test_t_gate():
    # reference test/integration/test_sampler_circuits.py:40-49
    m = _run_batches(K.t_gate_program(), 0, 100, 1)[0]
    assert np.count_nonzero(m) == 9


def test_r_gate_three_components():
    # reference test/integration/test_sampler_circuits.py:90-109
    m = _run_batches(K.three_coin_program(), 0, 10, 1)[0]
    assert [int(c) for c in m.sum(0)] == [7, 4, 0]


def test_exact_scalar_sum_reduces_while_adding():
    # reference test/unit/core/test_exact_scalar.py:66-83
    coeffs = np.array(
        [
            [[1, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0]],
            [[1, 0, 0, 0], [1, 0, 0, 0], [0, 2, 0, 0], [0, 2, 0, 0]],
        ]
    )
    powers = np.array([[0, 0, 0, 0], [3, 3, 2, 2]])
    s = ExactScalar.of(coeffs, powers).sum()
    assert np.array_equal(s.coeffs, [[1, 0, 0, 0], [1, 1, 0, 0]])
    assert np.array_equal(s.power, [2, 4])


def test_exact_scalar_mul_prod_sum_match_complex():
    # reference test/unit/core/test_exact_scalar.py (mul/prod/sum vs complex)
    rng = np.random.default_rng(0)
    a = rng.integers(-5, 6, size=(50, 4))
    b = rng.integers(-5, 6, size=(50, 4))
    w = np.exp(1j * np.pi / 4)
    basis = np.array([1, w, 1j, np.conj(w)])
    got = (ExactScalar.of(a) * ExactScalar.of(b)).to_complex()
    np.testing.assert_allclose(got, (a @ basis) * (b @ basis), rtol=1e-5, atol=1e-5)
    x = rng.integers(-3, 4, size=(10, 6, 4))
    p = ExactScalar.of(x).prod(axis=-1)
    np.testing.assert_allclose(p.to_complex(), np.prod(x @ basis, axis=-1), rtol=1e-4, atol=1e-3)
    powers = np.tile(np.arange(6), (10, 1))
    s = ExactScalar.of(x, powers).sum()
    np.testing.assert_allclose(s.to_complex(), np.sum((x @ basis) * 2.0**powers, axis=-1), rtol=1e-5, atol=1e-4)
    # canonical: not all coefficients even unless zero
    nz = np.any(s.coeffs != 0, axis=-1)
    assert np.all(np.any(s.coeffs[nz] & 1, axis=-1))


@pytest.mark.parametrize("P", (256, 300, 1024))
def test_matmul_gf2_no_uint8_saturation(P):
    # reference test/unit/utils/test_linalg.py:106-115
    a = np.ones((1, 1, P), np.uint8)
    b = np.ones((2, P), np.uint8)
    assert np.all(E.matmul_gf2(a, b) == P % 2)


def test_matmul_gf2_random_and_empty():
    # reference test/unit/utils/test_linalg.py:88-104,117-122
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2, (3, 4, 9)).astype(np.uint8)
    b = rng.integers(0, 2, (5, 9)).astype(np.uint8)
    want = (np.einsum("gtp,bp->bgt", a.astype(np.int64), b.astype(np.int64)) % 2).astype(np.uint8)
    assert np.array_equal(E.matmul_gf2(a, b), want)
    assert E.matmul_gf2(np.zeros((0, 3, 9), np.uint8), b).shape == (5, 0, 3)
    assert E.matmul_gf2(np.zeros((2, 0, 9), np.uint8), b).shape == (5, 2, 0)


# -- closed forms of reference test/unit/compile/test_terms.py:8-48 ------------------------------


def _ref_parity(bits, x):
    return ((x @ bits.reshape(-1, bits.shape[-1]).T) % 2).reshape(x.shape[0], bits.shape[0], bits.shape[1])


@pytest.mark.parametrize("seed", (0, 42))
def test_node_phases_closed_form(seed):
    np.random.seed(seed)
    G, T, P, B = 3, 4, 5, 7
    phases = np.random.randint(0, 8, size=(G, T)).astype(np.uint8)
    params = np.random.randint(0, 2, size=(G, T, P)).astype(np.uint8)
    counts = np.array([T, T - 1, 0], dtype=np.int32)
    x = np.random.randint(0, 2, size=(B, P)).astype(np.uint8)
    got = E.node_phases(NodePhases(phases, params, counts), x).to_complex()
    par = _ref_parity(params, x)
    term = 1 + np.exp(1j * np.pi * phases[None] / 4 + 1j * np.pi * par)
    mask = np.arange(T)[None, :] < counts[:, None]
    want = np.prod(np.where(mask[None], term, 1.0), axis=-1)
    np.testing.assert_allclose(got, want, atol=1e-6)


@pytest.mark.parametrize("seed", (0, 42))
def test_halfpi_closed_form(seed):
    np.random.seed(seed)
    G, T, P, B = 3, 4, 5, 7
    coeffs = np.random.choice([0, 2, 4, 6], size=(G, T)).astype(np.uint8)
    params = np.random.randint(0, 2, size=(G, T, P)).astype(np.uint8)
    x = np.random.randint(0, 2, size=(B, P)).astype(np.uint8)
    got = E.halfpi_phases(HalfPiPhases(coeffs, params), x).to_complex()
    want = np.prod(np.exp(1j * np.pi * coeffs[None] * _ref_parity(params, x) / 4), axis=-1)
    np.testing.assert_allclose(got, want, atol=1e-6)


@pytest.mark.parametrize("seed", (0, 42))
def test_pi_products_closed_form(seed):
    np.random.seed(seed)
    G, T, P, B = 3, 4, 5, 7
    pc = np.random.randint(0, 2, size=(G, T)).astype(np.uint8)
    pp = np.random.randint(0, 2, size=(G, T, P)).astype(np.uint8)
    fc = np.random.randint(0, 2, size=(G, T)).astype(np.uint8)
    fp = np.random.randint(0, 2, size=(G, T, P)).astype(np.uint8)
    x = np.random.randint(0, 2, size=(B, P)).astype(np.uint8)
    got = E.pi_products(PiProducts(pc, pp, fc, fp), x).to_complex()
    psi = (pc[None] + _ref_parity(pp, x)) % 2
    phi = (fc[None] + _ref_parity(fp, x)) % 2
    want = np.prod(np.exp(1j * np.pi * psi * phi), axis=-1)
    np.testing.assert_allclose(got, want, atol=1e-6)


@pytest.mark.parametrize("seed", (0, 42))
def test_phase_pairs_closed_form(seed):
    np.random.seed(seed)
    G, T, P, B = 3, 4, 5, 7
    al = np.random.randint(0, 8, size=(G, T)).astype(np.uint8)
    ap = np.random.randint(0, 2, size=(G, T, P)).astype(np.uint8)
    be = np.random.randint(0, 8, size=(G, T)).astype(np.uint8)
    bp = np.random.randint(0, 2, size=(G, T, P)).astype(np.uint8)
    counts = np.array([T, T - 1, 0], dtype=np.int32)
    x = np.random.randint(0, 2, size=(B, P)).astype(np.uint8)
    got = E.phase_pairs(PhasePairs(al, ap, be, bp, counts), x).to_complex()
    ea = np.exp(1j * np.pi * al[None] / 4 + 1j * np.pi * _ref_parity(ap, x))
    eb = np.exp(1j * np.pi * be[None] / 4 + 1j * np.pi * _ref_parity(bp, x))
    term = 1 + ea + eb - ea * eb
    mask = np.arange(T)[None, :] < counts[:, None]
    want = np.prod(np.where(mask[None], term, 1.0), axis=-1)
    np.testing.assert_allclose(got, want, atol=1e-5)


def test_families_with_zero_terms_are_identity():
    # reference test_terms.py:96-105,136-144,190-198,253-262
    G, P, B = 2, 3, 4
    x = np.zeros((B, P), np.uint8)
    z2, z3 = np.zeros((G, 0), np.uint8), np.zeros((G, 0, P), np.uint8)
    cnt = np.zeros(G, np.int32)
    for got in (
        E.node_phases(NodePhases(z2, z3, cnt), x),
        E.halfpi_phases(HalfPiPhases(z2, z3), x),
        E.pi_products(PiProducts(z2, z3, z2, z3), x),
        E.phase_pairs(PhasePairs(z2, z3, z2, z3, cnt), x),
    ):
        np.testing.assert_allclose(got.to_complex(), np.ones((B, G)))


def test_empty_program_evaluates_to_zero():
    # reference test/unit/compile/test_compile.py:31-46
    out = E.evaluate(empty_scalar_graphs(3), np.zeros((5, 3), np.uint8))
    assert out.dtype == np.complex64 and np.all(out == 0)


def test_complex_abs_edge_cases():
    re = np.array([0.0, 3.0, np.inf, np.nan, 0.0], np.float32)
    im = np.array([0.0, 4.0, np.inf, 1.0, -2.0], np.float32)
    got = E.complex_abs(re, im)
    assert got[0] == 0.0 and got[1] == 5.0 and np.isinf(got[2]) and got[4] == 2.0
