"""One process, several GPUs (``devices=``): bits equal the single-device run.  On a one-GPU box the same device is named
twice, which exercises the sharding, the thread pool and the shot offsets all the same.  `-m gpu`."""

import numpy as np
import pytest

import oracle
from tsim_b200.noise import ChannelSampler, DeviceChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program

pytestmark = pytest.mark.gpu


def _devices():
    from tsim_b200 import _lib

    n = _lib.load().tsb_device_count()
    return list(range(n)) if n > 1 else [0, 0]


@pytest.mark.parametrize("mode", ["auto", "fast"])
def test_multi_device_program_equals_single_device(mode):
    from tsim_b200.backend import DeviceProgram, MultiDeviceProgram

    prog = synthetic_program("cfg2_distill35")
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), 3e-3), seed=4).sample(20_001)
    one = DeviceProgram(prog, mode=mode)
    many = MultiDeviceProgram(prog, devices=_devices(), mode=mode)
    a, dev_a = one.sample(f, (3, 9))
    b, dev_b = many.sample(f, (3, 9))
    assert np.array_equal(a, b) and np.array_equal(dev_a, dev_b)
    ap, _ = one.sample(f, (3, 9), packed_out=True)
    bp, _ = many.sample(f, (3, 9), packed_out=True)
    assert np.array_equal(ap, bp)
    want = oracle.sample_program(prog, f[:1024], (3, 9), check_norm=False)
    assert np.array_equal(b[:1024], want)
    # a shard of a larger batch keeps its offsets
    c, dev_c = many.sample(f[5000:], (3, 9), shot_offset=5000)
    assert np.array_equal(c, a[5000:]) and not np.any(dev_c)
    x = np.random.default_rng(1).integers(0, 2, size=(777, one.level_params(0, 2))).astype(np.uint8)
    assert np.array_equal(one.evaluate(0, 2, x), many.evaluate(0, 2, x))


def test_sampler_class_with_devices(monkeypatch):
    import tsim_b200.sampler as S

    monkeypatch.setattr(S, "check_norm_deviations", lambda devs: None)
    prog = synthetic_program("cfg2_distill35")
    q = noise_probs(prog.infer_num_f(), 2e-3)
    devs = _devices()
    # host noise: the reference's exact stream, sharded over the devices
    a = S.CompiledDetectorSampler(prog, ChannelSampler.from_bit_probs(q, seed=1), seed=5).sample(30_000, batch_size=7000, append_observables=True)
    b = S.CompiledDetectorSampler(prog, ChannelSampler.from_bit_probs(q, seed=1), seed=5, devices=devs).sample(30_000, batch_size=7000, append_observables=True)
    assert np.array_equal(a, b)
    # device noise: K5 rows are a pure function of (seed, call, shot, channel), so the shards agree with one device
    a = S.CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(q, seed=1), seed=5).sample(30_000, batch_size=7000, bit_packed=True, use_detector_reference_sample=True)
    b = S.CompiledDetectorSampler(prog, DeviceChannelSampler.from_bit_probs(q, seed=1), seed=5, devices=devs).sample(30_000, batch_size=7000, bit_packed=True, use_detector_reference_sample=True)
    assert np.array_equal(a, b)
    # the seam: TSIM_B200_DEVICES routes sample_program (what install() rebinds) over the devices
    f = ChannelSampler.from_bit_probs(q, seed=2).sample(9000)
    want = S.sample_program(prog, f, (1, 2))
    monkeypatch.setenv("TSIM_B200_DEVICES", ",".join(str(d) for d in devs))
    got = S.sample_program(prog, f, (1, 2))
    assert np.array_equal(np.asarray(got), np.asarray(want))
