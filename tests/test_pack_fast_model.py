"""MODE_FAST packing + monoid arithmetic vs the oracle, on the CPU (no GPU needed)."""

import numpy as np
import pytest

import fast_model
from oracle import evaluation as E
from oracle.exact_scalar import ExactScalar
from tsim_b200 import pack as PK
from tsim_b200.program import CompiledComponent, make_program
from tsim_b200.synthetic import random_level


def _one_level_program(lv, F):
    # a 0-output component is not allowed; wrap as level 0 of a component with one trivial output level
    from kat_programs import const_level

    comp = CompiledComponent((0,), np.arange(F, dtype=np.int32), (lv, const_level(F + 1, -1)))
    return make_program([comp], num_f=max(F, 1))


@pytest.mark.parametrize("approx", [False, True])
@pytest.mark.parametrize("P,seed", [(5, 0), (31, 1), (40, 2), (70, 3)])
def test_fast_records_reproduce_oracle(P, seed, approx):
    rng = np.random.default_rng(seed)
    lv = random_level(rng, G=6, P=P, A=5, H=4, C=5, D=2, approx=approx, density=0.3)
    lv.prefactor.floatfactor[:] = rng.integers(-3, 4, size=(6, 4))
    lv.prefactor.floatfactor[0] = [1, 0, 0, 0]
    prog = _one_level_program(lv, P)
    pp = PK.pack_program(prog, mode="fast")
    xs = rng.integers(0, 2, size=(12, P)).astype(np.uint8)
    xs[0] = 0
    for x in xs:
        got = fast_model.evaluate_level(pp, 0, 0, x)
        if approx:
            re, im = E.evaluate_parts(lv, x[None, :])
            assert got[0] == "approx"
            assert np.float32(got[1]).tobytes() == re[0].tobytes() and np.float32(got[2]).tobytes() == im[0].tobytes()
        else:
            total = E.term_product(lv, x[None, :])
            with np.errstate(over="ignore"):
                s = ExactScalar(total.coeffs, (total.power + lv.prefactor.power2[None, :]).astype(np.int32)).sum()
            assert got[0] == "exact"
            if np.any(s.coeffs[0] != 0):
                assert np.array_equal(got[1], s.coeffs[0]) and got[2] == int(s.power[0])
            else:
                assert not np.any(got[1])


def test_auto_mode_falls_back_when_bound_fails():
    rng = np.random.default_rng(5)
    lv = random_level(rng, G=2, P=8, A=40, H=0, C=0, D=0, approx=False)
    lv.node_phases.counts[:] = 40
    prog = _one_level_program(lv, 8)
    ok, info = PK.reorder_is_exact(prog)
    assert not ok and info["log2_bound"] >= 40
    assert PK.pack_program(prog, mode="auto").mode == PK.MODE_FAITHFUL
    with pytest.raises(ValueError):
        PK.pack_program(prog, mode="fast")
