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
@pytest.mark.parametrize("density", [0.3, 0.06])  # 0.06: mostly one-word parities (the compact LIN_1 / PI_1 / PAIR_1 items)
@pytest.mark.parametrize("P,seed", [(5, 0), (31, 1), (40, 2), (70, 3)])
def test_sliced_records_reproduce_oracle(P, seed, approx, density):
    rng = np.random.default_rng(seed)
    lv = random_level(rng, G=6, P=P, A=5, H=4, C=5, D=4, approx=approx, density=density)
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


def test_sliced_chunk_layout_invariants():
    """Alignment and bounds the kernel relies on (16-byte items, zero entry in the record header, whole waves per chunk)."""
    from tsim_b200.pack_sliced import MAX_CHUNK_WORDS, SLICED_HEADER_WORDS
    from tsim_b200.synthetic import synthetic_program

    prog = synthetic_program("cfg2_distill35")
    pp = PK.pack_program(prog, mode="sliced")
    blob = pp.blob
    data = blob[int(blob[PK.H_OFF_DATA]) :]
    chunks = blob[int(blob[PK.H_OFF_CHUNK]) :][: int(blob[PK.H_N_CHUNKS]) * 4].reshape(-1, 4)
    levels = blob[int(blob[PK.H_OFF_LEVEL]) :][: int(blob[PK.H_N_LEVELS]) * PK.LEVEL_WORDS].reshape(-1, PK.LEVEL_WORDS)
    assert int(blob[PK.H_OFF_DATA]) % 32 == 0
    for lv in levels:
        first, n = int(lv[7]), int(lv[8])
        assert sum(int(c[2]) for c in chunks[first : first + n]) == int(lv[0])
        for ci, (off, words, ng, _) in enumerate(chunks[first : first + n]):
            off, words, ng = int(off), int(words), int(ng)
            assert off % 4 == 0 and words % 4 == 0 and words <= MAX_CHUNK_WORDS
            if ci + 1 < n:
                assert ng % 4 == 0  # only the last chunk of a level may end with a partial wave (waves of 4; of 8 on thin slices)
            chunk = data[off : off + words]
            end_of_records = None
            for g in range(ng):
                rec = int(chunk[g])
                assert rec % 4 == 0 and rec >= ((ng + 3) & ~3)
                hdr = chunk[rec : rec + SLICED_HEADER_WORDS]
                n_idx, nb = int(hdr[1]) & 0xFF, (int(hdr[1]) >> 8) & 0xFF
                assert 3 + nb <= n_idx <= 11 and not hdr[4:8].any()
                body = int(hdr[0]) & 0xFFFF
                assert body % 4 == 0 and (int(hdr[3]) & 0xFFFF) == SLICED_HEADER_WORDS + body
                main = int(hdr[3]) >> 16  # [main | aux]: the aux part holds whole pi runs, about half of the row loads
                assert main % 4 == 0 and 0 < main <= body
                tbl = int(hdr[2])
                assert tbl % 4 == 0 and tbl + 2 * (1 << n_idx) <= words
                end_of_records = rec + (int(hdr[3]) & 0xFFFF)
                assert tbl >= end_of_records or g + 1 < ng


def test_auto_mode_prefers_sliced_then_rowwise():
    from tsim_b200.synthetic import synthetic_program

    assert PK.pack_program(synthetic_program("cfg2_distill35")).mode == PK.MODE_SLICED
    assert PK.pack_program(synthetic_program("cfg2_distill35"), mode="rowwise").mode == PK.MODE_FAST
    # exact levels decode general phase pairs in two stages (table x ring factors), so cfg4's tables stay small enough
    pp4 = PK.pack_program(synthetic_program("cfg4_cultivation_d3"))
    assert pp4.mode == PK.MODE_SLICED and pp4.stats["data_bytes"] < 6 << 20 and int(pp4.blob[PK.H_PLANE_ROWS]) > 12
    # the size budget still sends bulky programs to the per-row records
    assert PK.pack_program(synthetic_program("cfg4_cultivation_d3"), sliced_budget_bytes=1 << 20).mode in (PK.MODE_FAST, PK.MODE_SLICED)
    # a program whose int32 arithmetic may wrap keeps the reference's operation order
    rng = np.random.default_rng(5)
    lv = random_level(rng, G=4, P=6, A=3, H=2, C=2, D=1, approx=False, density=0.4)
    lv.prefactor.floatfactor[:] = 1 << 29
    prog = _one_level_program(lv, 6)
    assert PK.pack_program(prog).mode == PK.MODE_FAITHFUL


def test_headline_program_all_levels_bit_identical_to_oracle():
    """The benchmark program (cfg2 shape, approximate branch): decode tables + plane arithmetic of every level."""
    from tsim_b200.synthetic import synthetic_program

    prog = synthetic_program("cfg2_distill35")
    pp = PK.pack_program(prog)
    assert pp.mode == PK.MODE_SLICED
    rng = np.random.default_rng(1)
    comp = prog.components[0]
    F = len(comp.f_selection)
    for k, lv in enumerate(comp.compiled_scalar_graphs):
        xs = (rng.random((24, F + k)) < 0.2).astype(np.uint8)
        xs[0] = 0
        xs[1] = 1
        got = sliced_model.evaluate_level(pp, 0, k, xs)
        re, im = E.evaluate_parts(lv, xs)
        for s, g in enumerate(got):
            assert np.float32(g[1]).tobytes() == re[s].tobytes() and np.float32(g[2]).tobytes() == im[s].tobytes(), (k, s)
