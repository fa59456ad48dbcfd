"""Full-size (BASELINE.json configuration) checks through size-independent properties.  `-m gpu`."""

import numpy as np
import pytest

import oracle
from tsim_b200.noise import ChannelSampler
from tsim_b200.synthetic import noise_probs, synthetic_program

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fast", "sliced", "sliced-direct"])  # "sliced" = with its default pattern cache
def cfg2(request):
    from tsim_b200.backend import DeviceProgram

    prog = synthetic_program("cfg2_distill35")
    return prog, DeviceProgram(prog, mode=request.param)


def test_million_shots_deterministic_sharded_and_fully_checked(cfg2):
    prog, dp = cfg2
    B = 1_000_000
    cs = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=12345)
    f = cs.sample_packed(B)
    key = (0, 7)
    a, dev_a = dp.sample(f, key, packed_out=True)
    b, dev_b = dp.sample(f, key, packed_out=True)
    assert np.array_equal(a, b) and np.array_equal(dev_a, dev_b)  # idempotent under a fixed key
    # shard invariance: 8 ragged shards with shot offsets == one batch (multi-GPU contract)
    cuts = [0, 1, 100_003, 250_000, 499_999, 500_512, 777_777, 999_999, B]
    parts = [dp.sample(f[lo:hi], key, shot_offset=lo, packed_out=True)[0] for lo, hi in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts), a)
    # every one of the 10^6 shots against the C restatement of the oracle (all host cores, a few seconds), and three
    # windows against the NumPy oracle itself (RNG counters = in-batch indices)
    from oracle import cport

    bits = np.unpackbits(a.view(np.uint8), axis=1, bitorder="little", count=prog.num_outputs).astype(bool)
    fb = np.unpackbits(f.view(np.uint8), axis=1, bitorder="little", count=prog.infer_num_f())
    want_all, want_dev = cport.sample_program(prog, fb, key, return_deviations=True, threads=cport.max_threads())
    assert np.array_equal(bits, want_all), f"{np.count_nonzero(bits != want_all)} differing bits in 10^6 shots"
    assert np.array_equal(np.asarray(dev_a, np.float32), np.asarray(want_dev, np.float32))
    for lo in (0, 314_159, B - 1024):
        want = oracle.sample_program(prog, fb[lo : lo + 1024], key, shot_offset=lo, check_norm=False)
        assert np.array_equal(bits[lo : lo + 1024], want)
    # direct columns are a pure function of f
    nd = len(prog.direct_f_indices)
    dest = np.empty(prog.num_outputs, np.int64)
    dest[prog.output_reindex] = np.arange(prog.num_outputs)
    assert np.array_equal(bits[:, dest[:nd]], fb[:, prog.direct_f_indices].astype(bool) ^ prog.direct_flips)
    # different key -> different compiled columns, same direct columns
    c, _ = dp.sample(f, (0, 8), packed_out=True)
    assert not np.array_equal(c, a)


def test_byte_and_packed_interfaces_agree_at_size(cfg2):
    prog, dp = cfg2
    B = 300_000
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=1).sample(B)
    from tsim_b200.noise import pack_f_rows

    bits, _ = dp.sample(f, (5, 5))
    packed, _ = dp.sample(pack_f_rows(f), (5, 5), packed_out=True)
    assert bits.shape == (B, prog.num_outputs)
    assert np.array_equal(np.packbits(bits, axis=1, bitorder="little"), packed.view(np.uint8)[:, : (prog.num_outputs + 7) // 8])


@pytest.mark.parametrize("mode", ["fast", "sliced", "sliced-direct"])
@pytest.mark.parametrize("name,B", [("cfg4_cultivation_d3", 4096), ("cfg5_distill85", 4096), ("cfg3_surface_d5", 200_000)])
def test_other_baseline_configs_match_oracle(name, B, mode):
    from tsim_b200.backend import DeviceProgram

    prog = synthetic_program(name)
    dp = DeviceProgram(prog, mode=mode)
    f = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=3).sample(B)
    got, dev = dp.sample(f, (1, 1))
    n = min(B, 1024)
    want, want_dev = oracle.sample_program(prog, f[:n], (1, 1), return_deviations=True, check_norm=False)
    assert np.array_equal(got[:n], want)
    assert np.array_equal(np.asarray(dev, np.float32), np.asarray(want_dev, np.float32))
    from oracle import cport

    full = cport.sample_program(prog, f, (1, 1), threads=cport.max_threads())  # the whole batch, C restatement
    assert np.array_equal(got, full)


@pytest.mark.parametrize("approx,B", [(True, 2_200_000), (False, 700_001)])
def test_sliced_multi_round_launches_match_per_row_kernel(approx, B):
    """Device-resident batches larger than one wave of slab groups (several rounds per CTA; 8-way split for the exact
    branch): the sliced kernel must reproduce the per-row kernel bit for bit (which the other tests pin to the oracle)."""
    import torch

    from test_gpu_parity import _random_program
    from tsim_b200.backend import DeviceProgram

    prog = synthetic_program("cfg2_distill35") if approx else _random_program(21, approx=False, n_comp=2, n_c=3, F=12, num_f=30, G=9)
    nf = prog.infer_num_f()
    cs = ChannelSampler.from_bit_probs(noise_probs(nf, 2e-3), seed=3)
    f = torch.from_numpy(cs.sample_packed(B).view(np.int64)).cuda()
    outs = {}
    for mode in ("sliced", "sliced-direct", "fast"):
        dp = DeviceProgram(prog, mode=mode)
        out = torch.zeros((B, dp.info["words_out64"]), dtype=torch.int64, device="cuda")
        dev = torch.zeros(max(1, dp.info["n_components"]), dtype=torch.float32, device="cuda")
        dp.sample_device(f.data_ptr(), B, (9, 1), out.data_ptr(), d_norm_dev=dev.data_ptr())
        torch.cuda.synchronize()
        outs[mode] = (out.cpu().numpy(), dev.cpu().numpy())
        if mode.startswith("sliced"):
            assert dp.info["mode"] == 2 and (dp.pattern_cache is None) == mode.endswith("-direct")
    for mode in ("sliced", "sliced-direct"):
        assert np.array_equal(outs[mode][0], outs["fast"][0])
        assert np.array_equal(outs[mode][1], outs["fast"][1])
    # and a window of it against the oracle
    lo = B - 2048
    fb = np.unpackbits(f[lo:].cpu().numpy().view(np.uint8), axis=1, bitorder="little", count=nf)
    want = oracle.sample_program(prog, fb, (9, 1), shot_offset=lo, check_norm=False)
    got = np.unpackbits(outs["sliced"][0][lo:].view(np.uint8), axis=1, bitorder="little", count=prog.num_outputs).astype(bool)
    assert np.array_equal(got, want)
