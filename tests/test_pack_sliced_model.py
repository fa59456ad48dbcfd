"""MODE_SLICED packing + bit-plane arithmetic vs the oracle, on the CPU."""

import numpy as np
import pytest

import sliced_model
from kat_programs import const_level
from oracle import evaluation as E
from oracle.exact_scalar import ExactScalar
from tsim_b200 import pack as PK
from tsim_b200.pack_sliced import MONOID, UNIT, _zw_mul, _zw_pow, monoid_exponents, pair_factor, ONE_PLUS_SQRT2, SQRT2_MINUS_ONE, ONE_PLUS_W
from tsim_b200.program import CompiledComponent, make_program
from tsim_b200.synthetic import random_level


def _one_level_program(lv, F):
    comp = CompiledComponent((0,), np.arange(F, dtype=np.int32), (lv, const_level(F + 1, -1)))
    return make_program([comp], num_f=max(F, 1))


def test_monoid_exponents_of_all_pair_factors():
    n_monoid = 0
    for al in range(8):
        for be in range(8):
            c = pair_factor(al, be)
            e = monoid_exponents(c)
            if e == "zero":
                assert not any(c)
            elif e is not None:
                a, b, n = e
                got = _zw_mul(_zw_mul(UNIT[a], _zw_pow(ONE_PLUS_SQRT2 if b >= 0 else SQRT2_MINUS_ONE, abs(b))), _zw_pow(ONE_PLUS_W, n))
                assert got == c
                n_monoid += 1
            else:
                assert al % 2 == 1 and be % 2 == 1  # only odd-odd factors contain other primes
    assert n_monoid >= 40
    for k, e in MONOID.items():
        assert monoid_exponents(tuple(int(i == 0) + UNIT[k][i] for i in range(4))) == e


@pytest.mark.parametrize("approx", [False, True])
@pytest.mark.parametrize("P,seed", [(5, 0), (31, 1), (40, 2), (70, 3)])
def test_sliced_records_reproduce_oracle(P, seed, approx):
    rng = np.random.default_rng(seed)
    lv = random_level(rng, G=6, P=P, A=5, H=4, C=5, D=4, approx=approx, density=0.3)
    lv.prefactor.floatfactor[:] = rng.integers(-3, 4, size=(6, 4))
    lv.prefactor.floatfactor[0] = [1, 0, 0, 0]
    prog = _one_level_program(lv, P)
    pp = PK.pack_program(prog, mode="sliced")
    assert pp.mode == PK.MODE_SLICED
    xs = rng.integers(0, 2, size=(40, P)).astype(np.uint8)
    xs[0] = 0
    got = sliced_model.evaluate_level(pp, 0, 0, xs)
    if approx:
        re, im = E.evaluate_parts(lv, xs)
        for s, g in enumerate(got):
            assert g[0] == "approx"
            assert np.float32(g[1]).tobytes() == re[s].tobytes() and np.float32(g[2]).tobytes() == im[s].tobytes()
    else:
        total = E.term_product(lv, xs)
        with np.errstate(over="ignore"):
            ssum = ExactScalar(total.coeffs, (total.power + lv.prefactor.power2[None, :]).astype(np.int32)).sum()
        for s, g in enumerate(got):
            assert g[0] == "exact"
            if np.any(ssum.coeffs[s] != 0):
                assert np.array_equal(g[1], ssum.coeffs[s]) and g[2] == int(ssum.power[s]), s
            else:
                assert not np.any(g[1])


@pytest.mark.parametrize("approx", [False, True])
def test_graphs_with_many_general_pairs_are_split_into_gated_variants(approx):
    rng = np.random.default_rng(7)
    P = 12
    lv = random_level(rng, G=3, P=P, A=3, H=2, C=2, D=6, approx=approx, density=0.3)
    lv.phase_pairs.alpha[:] = 2 * rng.integers(0, 4, size=lv.phase_pairs.alpha.shape) + 1  # odd-odd: never monoid-type
    lv.phase_pairs.beta[:] = 2 * rng.integers(0, 4, size=lv.phase_pairs.beta.shape) + 1
    lv.phase_pairs.counts[:] = [6, 5, 4]
    prog = _one_level_program(lv, P)
    pp = PK.pack_program(prog, mode="sliced")
    ch = pp.blob[int(pp.blob[PK.H_OFF_CHUNK]) :][: int(pp.blob[PK.H_N_CHUNKS]) * 4].reshape(-1, 4)
    assert int(ch[:, 2].sum()) > 3 + 2  # level 0 was split (the constant level 1 adds its own graphs)
    xs = rng.integers(0, 2, size=(64, P)).astype(np.uint8)
    got = sliced_model.evaluate_level(pp, 0, 0, xs)
    if approx:
        re, im = E.evaluate_parts(lv, xs)
        for s, g in enumerate(got):
            assert np.float32(g[1]).tobytes() == re[s].tobytes() and np.float32(g[2]).tobytes() == im[s].tobytes()
    else:
        total = E.term_product(lv, xs)
        with np.errstate(over="ignore"):
            ssum = ExactScalar(total.coeffs, (total.power + lv.prefactor.power2[None, :]).astype(np.int32)).sum()
        for s, g in enumerate(got):
            if np.any(ssum.coeffs[s] != 0):
                assert np.array_equal(g[1], ssum.coeffs[s]) and g[2] == int(ssum.power[s]), s
            else:
                assert not np.any(g[1])
