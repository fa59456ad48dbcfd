"""A program that travelled as ``.npz`` (the route real tsim programs take, tools/dump_tsim_programs.py) samples the same
bits on the device as the oracle computes from the original object; dumps with a golden batch from tsim itself
(``programs/*.golden.npz``, produced where tsim is installed) are compared with tsim's own bits.  `-m gpu`."""

import glob
import os

import numpy as np
import pytest

import oracle
from tsim_b200.noise import ChannelSampler
from tsim_b200.program import load_npz, load_npz_noise, save_npz
from tsim_b200.synthetic import noise_probs, synthetic_program

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["cfg2_distill35", "cfg4_cultivation_d3", "cfg3_surface_d5"])
def test_npz_round_trip_through_the_device(name, tmp_path):
    from tsim_b200.backend import DeviceProgram

    prog = synthetic_program(name)
    nf = prog.infer_num_f()
    cs = ChannelSampler.from_bit_probs(noise_probs(nf, 2e-3), seed=5)
    path = str(tmp_path / (name + ".npz"))
    save_npz(path, prog, noise=cs)
    back = load_npz(path)
    sparse, back_nf = load_npz_noise(path)
    f = ChannelSampler.from_sparse(sparse, back_nf, seed=5).sample(3000)
    assert np.array_equal(f, cs.sample(3000))
    got, dev = DeviceProgram(back).sample(f, (4, 2))
    want, want_dev = oracle.sample_program(prog, f, (4, 2), return_deviations=True, check_norm=False)
    assert np.array_equal(got, want)
    assert np.array_equal(np.asarray(dev, np.float32), np.asarray(want_dev, np.float32))


@pytest.mark.parametrize("golden", sorted(glob.glob(os.path.join(ROOT, "programs", "*.golden.npz"))) or [None])
def test_golden_batch_from_tsim(golden):
    if golden is None:
        pytest.skip("no tsim dump in programs/ (tools/dump_tsim_programs.py runs where tsim is installed)")
    from tsim_b200.backend import DeviceProgram

    z = np.load(golden)
    prog = load_npz(golden.replace(".golden.npz", ".npz"))
    f = np.unpackbits(z["f"], axis=1, bitorder="little", count=int(z["num_f"][0]))
    want = np.unpackbits(z["bits"], axis=1, bitorder="little", count=int(z["n_out"][0])).astype(bool)
    key = tuple(int(v) for v in z["key"].reshape(2))
    got, _ = DeviceProgram(prog).sample(f, key)
    assert np.array_equal(got, want), f"{np.count_nonzero((got != want).any(axis=1))} of {len(f)} shots differ from tsim's own bits"
    assert np.array_equal(oracle.sample_program(prog, f, key, check_norm=False), want)
