"""Randomised shapes: every kernel mode against the (C) oracle.  `-m gpu`."""

import shutil

import numpy as np
import pytest

import oracle
from tsim_b200.program import make_program
from tsim_b200.synthetic import synthetic_component

pytestmark = pytest.mark.gpu


def _oracle_bits(prog, f, key):
    if shutil.which("gcc") and shutil.which("make"):
        from oracle import cport

        return cport.sample_program(prog, f, key, return_deviations=True, threads=8)
    return oracle.sample_program(prog, f, key, return_deviations=True, check_norm=False)


def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    num_f = int(rng.integers(1, 230))
    n_comp = int(rng.integers(1, 4))
    n_direct = int(rng.integers(0, min(num_f, 70)))
    approx = bool(rng.integers(0, 2))
    comps, out = [], n_direct
    for _ in range(n_comp):
        F = int(rng.integers(0, min(num_f, 200) + 1))
        n_c = int(rng.integers(1, 5))
        G = [int(rng.integers(0 if k else 1, 7)) for k in range(n_c + 1)]  # some levels have no graph at all
        A, H, C, D = (int(rng.integers(0, 7)) for _ in range(4))
        comps.append(
            synthetic_component(rng, n_c, F, G, np.arange(num_f), out, A=A, H=H, C=C, D=D, approx=approx,
                                density=float(rng.uniform(0.02, 0.5)))
        )
        out += n_c
    prog = make_program(
        comps,
        direct_f_indices=rng.choice(num_f, n_direct, replace=False) if n_direct else [],
        direct_flips=rng.integers(0, 2, n_direct).astype(bool) if n_direct else [],
        output_order=rng.permutation(out),
        num_f=num_f,
    )
    B = int(rng.integers(1, 700))
    f = (rng.random((B, num_f)) < rng.uniform(0.0, 0.3)).astype(np.uint8)
    key = (int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32)))
    return prog, f, key


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_all_modes(seed):
    from tsim_b200.backend import DeviceProgram

    prog, f, key = _random_case(seed)
    want, want_dev = _oracle_bits(prog, f, key)
    wd = np.asarray(want_dev, np.float32).view(np.uint32)
    for mode in ("faithful", "fast", "sliced", "sliced-direct"):
        try:
            dp = DeviceProgram(prog, mode=mode)
        except ValueError:
            assert mode != "faithful"  # reordered modes may decline a program, the faithful one never does
            continue
        got, dev = dp.sample(f, key)
        assert np.array_equal(got, want), (seed, mode, int(np.count_nonzero(got != want)))
        assert np.array_equal(np.asarray(dev, np.float32).view(np.uint32), wd), (seed, mode)
