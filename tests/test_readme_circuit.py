"""BASELINE.json config 0: the README 2-qubit circuit (reference README.md:40-55), 100 shots.

    RX 0; R 1; T 0; PAULI_CHANNEL_1(0.1, 0.1, 0.2) 0 1; H 0; CNOT 0 1; DEPOLARIZE2(0.01) 0 1; M 0 1; DETECTOR rec[-1] rec[-2]

The state before measurement is a|00> + b|11>, so the detector m0 ^ m1 is deterministic: tsim classifies it as a
*direct* output and the whole program is plumbing -- one f bit, no compiled component.  Propagating the
Paulis by hand: errors on qubit 0 before H never flip the detector; X or Y on qubit 1 before the CNOT do (0.1 + 0.1);
of the 15 two-qubit Paulis after the CNOT the 8 with an X/Y on exactly one qubit do (8 * 0.01 / 15).
"""

import numpy as np
import pytest

from tsim_b200.noise import ChannelSampler
from tsim_b200.program import make_program

P_Q1 = 0.2
P_DEP = 8 * 0.01 / 15
P_DET = P_Q1 * (1 - P_DEP) + (1 - P_Q1) * P_DEP


def _readme_program():
    return make_program([], direct_f_indices=[0], direct_flips=[False], num_outputs=1, num_detectors=1, num_f=1)


def _readme_channels(seed):
    one = np.array([[1]], dtype=np.uint8)
    return ChannelSampler.from_sparse([(P_Q1, np.array([1.0]), one), (P_DEP, np.array([1.0]), one)], 1, seed=seed)


def test_readme_detector_rate_host_noise_oracle():
    import oracle

    prog = _readme_program()
    f = _readme_channels(1).sample(200_000)
    bits = oracle.sample_program(prog, f, (0, 0))
    assert bits.shape == (200_000, 1)
    assert abs(bits.mean() - P_DET) < 4 * np.sqrt(P_DET * (1 - P_DET) / 200_000)


@pytest.mark.gpu
def test_readme_circuit_through_the_detector_sampler():
    from tsim_b200.noise import DeviceChannelSampler
    from tsim_b200.sampler import CompiledDetectorSampler

    prog = _readme_program()
    s = CompiledDetectorSampler(prog, _readme_channels(3), seed=0)
    out = s.sample(shots=100)
    assert out.shape == (100, 1) and out.dtype == np.bool_
    big = CompiledDetectorSampler(prog, _readme_channels(3), seed=0).sample(shots=200_000)
    assert abs(big.mean() - P_DET) < 4 * np.sqrt(P_DET * (1 - P_DET) / 200_000)
    # device program + device noise on the same circuit (the direct path through K1/K5)
    from tsim_b200.backend import DeviceProgram

    dp = DeviceProgram(prog)
    noise = DeviceChannelSampler(_readme_channels(0)._sparse_data, 1, seed=5)
    bits, dev = dp.sample_noisy(noise, 200_000, (0, 1))
    assert len(dev) == 0 and abs(bits.mean() - P_DET) < 4 * np.sqrt(P_DET * (1 - P_DET) / 200_000)
