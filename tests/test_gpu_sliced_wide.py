"""The wide layout of the sliced kernel (64-bit lanes, two slabs per lane; sliced_kernels.cuh) against the oracle.

The wide layout is opt-in (TSIM_B200_SLICED_WIDE=1: measured slower than the narrow one on the headline workload, see
DESIGN.md); the tuning knobs force it here so that the cases the layout adds are covered at sizes the oracle finishes
in seconds: full groups, half-populated groups (lanes 16..31 idle), ragged tails inside a unit, several rounds, the
row-list variant behind the pattern cache, and the programs that must stay on the narrow layout (exact levels, more
than 127 rows).  `-m gpu`."""

import numpy as np
import pytest

import oracle
from oracle import cport
from tsim_b200 import pack as PK
from tsim_b200.noise import ChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program
from test_gpu_parity import _dev_equal, _device_program, _oracle, _random_program

pytestmark = pytest.mark.gpu


@pytest.fixture
def wide(monkeypatch):
    monkeypatch.setenv("TSIM_B200_SLICED_SPLIT", "4")
    monkeypatch.setenv("TSIM_B200_SLICED_WIDE", "1")


@pytest.mark.parametrize("mode", ["sliced", "sliced-direct"])
@pytest.mark.parametrize("B", [1, 700, 1024, 1500, 2048, 3000, 5000])  # units: 1 (half group), 1, 1, 2, 2, 3 (full + half), 5
@pytest.mark.parametrize("shape", [dict(F=10, num_f=24), dict(F=40, num_f=70, n_c=4)])
def test_wide_layout_matches_oracle(wide, shape, B, mode):
    prog = _random_program(17, approx=True, **shape)
    f = (np.random.default_rng(5).random((B, prog.infer_num_f())) < 0.1).astype(np.uint8)
    f[0] = 0
    key = (21, 43)
    want, want_dev = cport.sample_program(prog, f, key, return_deviations=True)
    dp = _device_program(prog, mode)
    assert int(PK.pack_program(prog, mode="sliced").blob[PK.H_INDEX_SCALE]) == 2
    got, dev = dp.sample(f, key)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} differing bits"
    assert _dev_equal(dev, want_dev)
    # a shot offset moves the RNG counters, not the layout
    got2, _ = dp.sample(f[B // 2 :], key, shot_offset=B // 2)
    assert np.array_equal(got2, want[B // 2 :])


@pytest.mark.parametrize("mode", ["sliced", "sliced-direct"])
def test_wide_layout_several_rounds(wide, mode, monkeypatch):
    # a tight shared-memory limit leaves room for one wide group: nine units go through in rounds
    monkeypatch.setenv("TSIM_B200_SMEM_LIMIT", str(120 * 1024))
    prog = synthetic_program("cfg2_distill35")
    B = 148 * 1024 * 2 + 9000  # more units than SMs
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=99).sample(B)
    key = (3, 4)
    want, want_dev = cport.sample_program(prog, f, key, return_deviations=True, threads=cport.max_threads())
    got, dev = _device_program(prog, mode).sample(f, key)
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} differing bits"
    assert _dev_equal(dev, want_dev)


def test_wide_equals_narrow_on_a_chip_filling_batch(monkeypatch):
    prog = synthetic_program("cfg2_distill35")
    B = 148 * 7 * 1024 - 5000
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=7).sample_packed(B)
    key = (0, 11)
    dp = _device_program(prog, "sliced-direct")
    a, dev_a = dp.sample(f, key, packed_out=True)  # plan's own choice
    monkeypatch.setenv("TSIM_B200_SLICED_WIDE", "0")
    b, dev_b = dp.sample(f, key, packed_out=True)
    monkeypatch.setenv("TSIM_B200_SLICED_WIDE", "1")
    c, dev_c = dp.sample(f, key, packed_out=True)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert _dev_equal(dev_a, dev_b) and _dev_equal(dev_a, dev_c)


@pytest.mark.parametrize("shape,approx", [(dict(F=10, num_f=24), False), (dict(F=150, num_f=200, n_c=2, n_comp=1, G=3), True)])
def test_programs_that_stay_narrow(wide, shape, approx):
    # exact levels (four-word accumulators) and programs with more than 127 rows (plain index bytes) ignore the knob
    prog = _random_program(23, approx=approx, **shape)
    B = 3000
    f = (np.random.default_rng(8).random((B, prog.infer_num_f())) < 0.1).astype(np.uint8)
    key = (9, 9)
    want, want_dev = _oracle(prog, f, key)
    got, dev = _device_program(prog, "sliced-direct").sample(f, key)
    assert np.array_equal(got, want) and _dev_equal(dev, want_dev)
    if approx:
        assert int(PK.pack_program(prog, mode="sliced").blob[PK.H_INDEX_SCALE]) == 1
