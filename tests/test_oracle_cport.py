"""The C restatement (oracle/c/oracle.c) must agree bit for bit with the NumPy oracle (CPU only)."""

import shutil

import numpy as np
import pytest

import oracle
from tsim_b200.noise import ChannelSampler
from tsim_b200.program import CompiledComponent, make_program
from tsim_b200.synthetic import noise_probs, random_level, synthetic_program

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None or shutil.which("make") is None, reason="needs gcc + make")


def _same(prog, f, key, **kw):
    from oracle import cport

    want, wd = oracle.sample_program(prog, f, key, return_deviations=True, check_norm=False, **kw)
    got, gd = cport.sample_program(prog, f, key, return_deviations=True, threads=4, **kw)
    assert np.array_equal(got, want)
    assert np.array_equal(np.asarray(gd, np.float32).view(np.uint32), np.asarray(wd, np.float32).view(np.uint32))


@pytest.mark.parametrize("name,B", [("cfg2_distill35", 1500), ("cfg3p_rank1", 600), ("cfg5_distill85", 300), ("cfg3_surface_d5", 500)])
def test_c_port_matches_numpy_oracle(name, B):
    prog = synthetic_program(name)
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), 4e-3), seed=5).sample(B)
    _same(prog, f, (3, 1))
    _same(prog, f[100:400], (3, 1), shot_offset=100)


def test_c_port_wraps_like_numpy():
    rng = np.random.default_rng(77)
    F = 6
    levels = []
    for k in range(3):
        lv = random_level(rng, G=3, P=F + k, A=44, H=2, C=2, D=3, approx=False, density=0.4, power2_range=(-3, 3))
        lv.node_phases.counts[:] = 44
        lv.node_phases.phases[:] = rng.choice([1, 3, 5, 7], size=lv.node_phases.phases.shape)
        lv.prefactor.floatfactor[:] = rng.integers(-9, 10, size=(3, 4))
        levels.append(lv)
    prog = make_program([CompiledComponent((0, 1), np.arange(F, dtype=np.int32), tuple(levels))], num_f=F)
    _same(prog, rng.integers(0, 2, size=(300, F)).astype(np.uint8), (6, 6))
