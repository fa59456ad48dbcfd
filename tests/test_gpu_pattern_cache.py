"""The optional pattern cache must not change a single bit.  `-m gpu`."""

import numpy as np
import pytest

import oracle
from tsim_b200.noise import ChannelSampler, DeviceChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fast", "faithful", "sliced"])
@pytest.mark.parametrize("name,B,p", [("cfg2_distill35", 300_000, 1e-3), ("cfg2_distill35", 40_000, 2e-2), ("cfg3p_rank1", 20_000, 5e-3), ("cfg4_cultivation_d3", 20_000, 1e-3)])
def test_cache_is_bit_identical_to_full_evaluation(name, B, p, mode):
    from tsim_b200.backend import DeviceProgram

    prog = synthetic_program(name)
    try:
        dp = DeviceProgram(prog, mode=mode, pattern_cache=None)
    except ValueError:
        pytest.skip("no sliced records for this program")
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f(), p), seed=21).sample_packed(B)
    key = (3, 14)
    base, base_dev = dp.sample(f, key, packed_out=True)
    base = base.copy()
    for wmax in (0, 1, 2, 3):
        n = dp.set_pattern_cache(wmax)
        assert n > 0
        got, dev = dp.sample(f, key, packed_out=True)
        assert np.array_equal(got, base), f"wmax={wmax}: {np.count_nonzero(got != base)} rows differ"
        assert np.array_equal(np.asarray(dev, np.float32).view(np.uint32), np.asarray(base_dev, np.float32).view(np.uint32))
    # sharded with offsets, cache on
    cut = B // 3 + 5
    a, _ = dp.sample(f[:cut], key, packed_out=True)
    b, dev_b = dp.sample(f[cut:], key, shot_offset=cut, packed_out=True)
    assert np.array_equal(np.concatenate([a, b]), base) and not np.any(dev_b)
    dp.set_pattern_cache(None)
    again, _ = dp.sample(f, key, packed_out=True)
    assert np.array_equal(again, base)


def test_cache_against_oracle_and_on_fused_and_device_paths():
    import torch

    from tsim_b200.backend import DeviceProgram

    prog = synthetic_program("cfg2_distill35")
    dp = DeviceProgram(prog)
    dp.set_pattern_cache(2)
    nf = prog.infer_num_f()
    f = ChannelSampler.from_bit_probs(noise_probs(nf, 3e-3), seed=8).sample(4096)
    got, dev = dp.sample(f, (1, 5))
    want, want_dev = oracle.sample_program(prog, f, (1, 5), return_deviations=True, check_norm=False)
    assert np.array_equal(got, want) and np.array_equal(np.asarray(dev, np.float32), np.asarray(want_dev, np.float32))
    # fused noise -> sample
    noise = DeviceChannelSampler.from_bit_probs(noise_probs(nf, 1e-3), seed=4)
    bits, _, fp = dp.sample_noisy(noise, 150_000, (2, 2), return_f=True)
    fb = np.unpackbits(fp.view(np.uint8), axis=1, bitorder="little", count=nf)
    want = oracle.sample_program(prog, fb[:2048], (2, 2), check_norm=False)
    assert np.array_equal(bits[:2048], want)
    # device-pointer entry
    from tsim_b200.noise import pack_f_rows

    d_f = torch.from_numpy(pack_f_rows(f).view(np.int64)).cuda()
    d_out = torch.empty((4096, dp.info["words_out64"]), dtype=torch.int64, device="cuda")
    dp.sample_device(d_f.data_ptr(), 4096, (1, 5), d_out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = np.unpackbits(d_out.cpu().numpy().view(np.uint8), axis=1, bitorder="little", count=prog.num_outputs).astype(bool)
    assert np.array_equal(out, got)
