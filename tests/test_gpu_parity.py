"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a GPU: `-m gpu`."""

import os
import warnings

import numpy as np
import pytest

import kat_programs as K
import oracle
from oracle import evaluation as E
from tsim_b200.noise import ChannelSampler, pack_f_rows
from tsim_b200.program import CompiledComponent, make_program
from tsim_b200.synthetic import noise_probs, random_level, synthetic_component, synthetic_program

pytestmark = pytest.mark.gpu

MODES = ("faithful", "fast", "sliced", "sliced-direct")  # "sliced" runs with its default pattern cache, "-direct" without


def _device_program(prog, mode, **kw):
    from tsim_b200.backend import DeviceProgram

    return DeviceProgram(prog, mode=mode, **kw)


def _oracle(prog, f, key, shot_offset=0):
    return oracle.sample_program(prog, f, key, shot_offset=shot_offset, return_deviations=True, check_norm=False)


@pytest.fixture
def no_norm_check(monkeypatch):
    """Random programs are not probability trees: silence the ValueError / warning of sampler.py:149-161."""
    import tsim_b200.sampler as S

    monkeypatch.setattr(S, "check_norm_deviations", lambda devs: None)


def _dev_equal(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a, b)


def _random_program(seed, *, n_comp=2, n_c=3, F=10, num_f=24, G=5, A=4, H=3, C=4, D=2, approx=False, n_direct=3, density=0.3):
    rng = np.random.default_rng(seed)
    comps = []
    out = n_direct
    for _ in range(n_comp):
        comps.append(
            synthetic_component(rng, n_c, F, [G] * (n_c + 1), np.arange(num_f), out, A=A, H=H, C=C, D=D, approx=approx, density=density)
        )
        out += n_c
    return make_program(
        comps,
        direct_f_indices=rng.choice(num_f, n_direct, replace=False),
        direct_flips=rng.integers(0, 2, n_direct).astype(bool),
        output_order=rng.permutation(out),
        num_f=num_f,
    )


# ---------------------------------------------------------------------------------------------
# reference known-answer tests, through the sampler classes
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("mode", MODES)
def test_seed_chain_hm(mode):
    # reference test/unit/test_sampler.py:223-233
    from tsim_b200.sampler import CompiledMeasurementSampler

    for _ in range(2):
        s = CompiledMeasurementSampler(K.hm_program(), ChannelSampler.from_bit_probs([], seed=0), seed=0, mode=mode)
        assert [int(np.count_nonzero(s.sample(100))) for _ in range(4)] == [48, 53, 52, 50]


@pytest.mark.parametrize("mode", MODES)
def test_bell_t_gate_r_gate(mode):
    # reference test/integration/test_sampler_circuits.py:10-22, 40-49, 90-109
    from tsim_b200.sampler import CompiledMeasurementSampler

    nz = ChannelSampler.from_bit_probs([], seed=0)
    m = CompiledMeasurementSampler(K.bell_program(), nz, seed=0, mode=mode).sample(100)
    assert m.dtype == np.bool_ and np.array_equal(m[:, 0], m[:, 1]) and np.count_nonzero(m[:, 0]) == 48
    m = CompiledMeasurementSampler(K.t_gate_program(), nz, seed=0, mode=mode).sample(100)
    assert np.count_nonzero(m) == 9
    m = CompiledMeasurementSampler(K.three_coin_program(), ChannelSampler.from_bit_probs([0.0], seed=0), seed=0, mode=mode).sample(10)
    assert [int(c) for c in m.sum(0)] == [7, 4, 0]


# ---------------------------------------------------------------------------------------------
# bit-exact parity with the oracle on seeded random programs
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("approx", [False, True])
@pytest.mark.parametrize(
    "shape",
    [
        dict(F=10, num_f=24),  # W = 1
        dict(F=40, num_f=70, n_c=4),  # W = 2
        dict(F=90, num_f=130, n_c=2, n_comp=1),  # W = 3
        dict(F=150, num_f=200, n_c=2, n_comp=1, G=3),  # W = 5
    ],
)
def test_random_programs_match_oracle(shape, approx, mode):
    prog = _random_program(11, approx=approx, **shape)
    B = 1500  # three tiles, last one ragged
    f = (np.random.default_rng(3).random((B, prog.infer_num_f())) < 0.1).astype(np.uint8)
    f[0] = 0
    key = (123, 456)
    want, want_dev = _oracle(prog, f, key)
    dp = _device_program(prog, mode)
    got, dev = dp.sample(f, key)
    assert got.dtype == np.bool_ and got.shape == want.shape
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} differing bits"
    assert _dev_equal(dev, want_dev)


@pytest.mark.parametrize("mode", MODES)
def test_cfg2_shape_matches_oracle_resident_and_streamed(mode, monkeypatch):
    prog = synthetic_program("cfg2_distill35")
    B = 2048
    cs = ChannelSampler.from_bit_probs(noise_probs(prog.infer_num_f()), seed=12345)
    f = cs.sample(B)
    key = oracle.split((0, 0))[1]
    want, want_dev = _oracle(prog, f, key)
    dp = _device_program(prog, mode)
    assert dp.info["resident"] == (0 if mode.startswith("sliced") else 1)  # sliced CTAs spend their shared memory on per-shot state
    got, dev = dp.sample(f, key)
    assert np.array_equal(got, want) and _dev_equal(dev, want_dev)
    # packed formats
    gotp, _ = dp.sample(pack_f_rows(f), key, packed_out=True)
    assert np.array_equal(np.unpackbits(gotp.view(np.uint8), axis=1, bitorder="little", count=prog.num_outputs).astype(bool), want)
    # same program through the streamed path (shared memory capped -> chunk ring)
    monkeypatch.setenv("TSIM_B200_SMEM_LIMIT", str((140 if mode.startswith("sliced") else 64) * 1024))
    dps = _device_program(prog, mode)
    assert dps.info["resident"] == 0 and (not mode.startswith("sliced") or dps.info["threads"] < dp.info["threads"])
    got2, dev2 = dps.sample(f, key)
    assert np.array_equal(got2, want) and _dev_equal(dev2, want_dev)


@pytest.mark.parametrize("mode", MODES)
def test_many_small_components_cfg3p(mode):
    prog = synthetic_program("cfg3p_rank1")
    B = 700
    f = ChannelSampler.from_bit_probs(noise_probs(121, 5e-3), seed=1).sample(B)
    want, want_dev = _oracle(prog, f, (7, 7))
    got, dev = _device_program(prog, mode).sample(f, (7, 7))
    assert np.array_equal(got, want) and _dev_equal(dev, want_dev)


def test_direct_only_program_cfg3():
    prog = synthetic_program("cfg3_surface_d5")
    B = 1000
    f = ChannelSampler.from_bit_probs(noise_probs(121, 1e-2), seed=2).sample(B)
    want = oracle.sample_program(prog, f, (0, 0), check_norm=False)
    got, dev = _device_program(prog, "auto").sample(f, (0, 0))
    assert np.array_equal(got, want) and len(dev) == 0


@pytest.mark.parametrize("mode", MODES)
def test_shot_offset_shards_equal_full_batch(mode):
    prog = _random_program(5, approx=True)
    B = 1030
    f = (np.random.default_rng(1).random((B, prog.infer_num_f())) < 0.2).astype(np.uint8)
    dp = _device_program(prog, mode)
    full, dev_full = dp.sample(f, (9, 9))
    cut = 517
    a, dev_a = dp.sample(f[:cut], (9, 9), shot_offset=0)
    b, dev_b = dp.sample(f[cut:], (9, 9), shot_offset=cut)
    assert np.array_equal(np.concatenate([a, b]), full)
    assert _dev_equal(dev_a, dev_full) and not np.any(dev_b)
    want_b, _ = _oracle(prog, f[cut:], (9, 9), shot_offset=cut)
    assert np.array_equal(b, want_b)


@pytest.mark.parametrize("mode", MODES)
def test_edge_shapes(mode):
    prog = _random_program(2)
    dp = _device_program(prog, mode)
    nf = prog.infer_num_f()
    got, dev = dp.sample(np.zeros((0, nf), np.uint8), (1, 2))
    assert got.shape == (0, prog.num_outputs)
    f1 = np.ones((1, nf), np.uint8)
    want, want_dev = _oracle(prog, f1, (1, 2))
    got, dev = dp.sample(f1, (1, 2))
    assert np.array_equal(got, want) and _dev_equal(dev, want_dev)
    # level with zero graphs -> amplitude 0 (reference test_compile.py:31-46): p1 = 0 -> bit 0, prev stays
    from tsim_b200.program import empty_scalar_graphs

    comp = CompiledComponent((0,), np.zeros(0, np.int32), (K.const_level(0, 0), empty_scalar_graphs(1)))
    prog0 = make_program([comp], num_f=0)
    want, want_dev = _oracle(prog0, np.zeros((40, 0), np.uint8), (3, 3))
    got, dev = _device_program(prog0, mode).sample(np.zeros((40, 0), np.uint8), (3, 3))
    assert np.array_equal(got, want) and not got.any() and _dev_equal(dev, want_dev)


def test_bad_arguments_raise():
    from tsim_b200 import _lib
    import ctypes as C

    prog = _random_program(2)
    dp = _device_program(prog, "auto")
    with pytest.raises(ValueError):
        dp.sample(np.zeros((4, prog.infer_num_f() + 1), np.uint8), (0, 0))
    with pytest.raises(ValueError):
        dp.evaluate(99, 0, np.zeros((1, 1), np.uint8))
    bad = np.zeros(64, np.uint32)
    h = C.c_void_p()
    rc = _lib.load().tsb_program_create(bad.ctypes.data_as(C.c_void_p), bad.size, 0, C.byref(h))
    assert rc == -1 and b"magic" in _lib.load().tsb_last_error()


# ---------------------------------------------------------------------------------------------
# evaluate (marginals): float32 parts bit-identical to the oracle, hence within 1e-6 relative
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("approx", [False, True])
def test_evaluate_matches_oracle(mode, approx):
    prog = _random_program(21, approx=approx, F=40, num_f=70, n_c=3, n_comp=2)
    dp = _device_program(prog, mode)
    rng = np.random.default_rng(0)
    for ci, comp in enumerate(prog.components):
        for li, lv in enumerate(comp.compiled_scalar_graphs):
            x = rng.integers(0, 2, size=(300, lv.n_params)).astype(np.uint8)
            want = E.evaluate(lv, x)
            got = dp.evaluate(ci, li, x)
            assert got.dtype == np.complex64
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
            # tolerance north_star states for marginals: 1e-6 relative
            np.testing.assert_allclose(np.abs(got), np.abs(want), rtol=1e-6, atol=0)


def test_probability_of_joint_mode(no_norm_check):
    # joint-mode program (reference sampler.py:906-953): P(state | f) = |E_joint| / |E_norm| per component
    from tsim_b200.sampler import CompiledStateProbs

    rng = np.random.default_rng(4)
    F, n = 6, 3
    joint = random_level(rng, 4, F + n, 3, 2, 2, 1, approx=False, density=0.3)
    norm = random_level(rng, 2, F, 2, 1, 1, 0, approx=False, density=0.3)
    comp = CompiledComponent((1, 2, 3), np.arange(F, dtype=np.int32), (norm, joint))
    prog = make_program([comp], direct_f_indices=[6], direct_flips=[True], num_outputs=4, num_f=7)
    state = np.array([1, 0, 1, 1], dtype=np.uint8)
    sp = CompiledStateProbs(prog, ChannelSampler.from_bit_probs([0.2] * 7, seed=3), seed=0)
    got = sp.probability_of(state, batch_size=64)
    f = ChannelSampler.from_bit_probs([0.2] * 7, seed=3).sample(64)
    direct_ok = ((f[:, 6].astype(bool) ^ True) == bool(state[0])).astype(np.float32)
    fs = f[:, :F]
    pn = E.evaluate_abs(norm, fs)
    pj = E.evaluate_abs(joint, np.hstack([fs, np.tile(state[1:], (64, 1))]))
    with np.errstate(all="ignore"):
        want = direct_ok * pj / pn
    np.testing.assert_allclose(got, want, rtol=1e-6, equal_nan=True)


# ---------------------------------------------------------------------------------------------
# sampler plumbing on the device path
# ---------------------------------------------------------------------------------------------


def test_detector_sampler_layout_flags_and_reference_sample(no_norm_check):
    from tsim_b200.sampler import CompiledDetectorSampler

    prog = _random_program(8, n_comp=2, n_c=2, n_direct=3)
    prog.num_detectors = 5
    q = np.full(prog.infer_num_f(), 0.05)

    def mk():
        return CompiledDetectorSampler(prog, ChannelSampler.from_bit_probs(q, seed=4), seed=11)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        base = mk().sample(50, batch_size=16, append_observables=True)
        assert base.shape == (50, 7)
        det, obs = mk().sample(50, batch_size=16, separate_observables=True)
        assert np.array_equal(det, base[:, :5]) and np.array_equal(obs, base[:, 5:])
        assert np.array_equal(mk().sample(50, batch_size=16), base[:, :5])
        pre = mk().sample(50, batch_size=16, prepend_observables=True)
        assert np.array_equal(pre, np.concatenate([base[:, 5:], base[:, :5]], axis=1))
        packed = mk().sample(50, batch_size=16, append_observables=True, bit_packed=True)
        assert np.array_equal(packed, np.packbits(base, axis=1, bitorder="little"))
        # every layout, packed end to end == packbits of the bool result
        for kw in (dict(), dict(prepend_observables=True), dict(prepend_observables=True, append_observables=True),
                   dict(use_detector_reference_sample=True, append_observables=True),
                   dict(use_observable_reference_sample=True, prepend_observables=True)):
            a = mk().sample(50, batch_size=16, **kw)
            b = mk().sample(50, batch_size=16, bit_packed=True, **kw)
            assert np.array_equal(b, np.packbits(a, axis=1, bitorder="little")), kw
        d1, o1 = mk().sample(50, batch_size=16, separate_observables=True, use_detector_reference_sample=True)
        d2, o2 = mk().sample(50, batch_size=16, separate_observables=True, use_detector_reference_sample=True, bit_packed=True)
        assert np.array_equal(d2, np.packbits(d1, axis=1, bitorder="little")) and np.array_equal(o2, np.packbits(o1, axis=1, bitorder="little"))
        with pytest.raises(ValueError):
            mk().sample(5, separate_observables=True, append_observables=True)
        with pytest.raises(ValueError):
            mk().sample(-1)
        with pytest.raises(ValueError):
            mk().sample(5, batch_size=0)
        assert mk().sample(0).shape == (0, 5)
        # reference sample: oracle replay of the same schedule (batch bumped by one, row 0 zeroed)
        s = mk()
        got = s.sample(48, batch_size=16, append_observables=True, use_detector_reference_sample=True)
        cs = ChannelSampler.from_bit_probs(q, seed=4)
        key = (0, 11)
        rows, ref = [], None
        for _ in range(3):
            f = cs.sample(17)
            if ref is None:
                f[0] = 0
            key, sub = oracle.split(key)
            o = oracle.sample_program(prog, f, sub, check_norm=False)
            if ref is None:
                ref, o = o[0].copy(), o[1:]
            rows.append(o)
        want = np.concatenate(rows)[:48].copy()
        want[:, :5] ^= ref[:5]
        assert np.array_equal(got, want)


def test_postselection_matches_unmasked_rows(no_norm_check):
    from tsim_b200.sampler import CompiledDetectorSampler

    prog = _random_program(9, n_comp=1, n_c=3, n_direct=4)
    prog.num_detectors = 6
    q = np.full(prog.infer_num_f(), 0.1)
    mask = np.zeros(6, bool)
    direct_cols = np.asarray(prog.output_order[:4])
    mask[[c for c in direct_cols if c < 6][:2]] = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = CompiledDetectorSampler(prog, ChannelSampler.from_bit_probs(q, seed=6), seed=2)
        got = s.sample(200, batch_size=32, append_observables=True, postselection_mask=mask)
    assert got.shape == (200, 7)
    discarded = (got[:, :6] & mask).any(axis=1)
    assert discarded.any() and (~discarded).any()
    comp_cols = np.setdiff1d(np.arange(7), direct_cols)
    assert not got[discarded][:, comp_cols].any()


# ---------------------------------------------------------------------------------------------
# int32 wrap-around: the reference's arithmetic wraps silently (exact_scalar.py:31-39); MODE_FAITHFUL must wrap the
# same way, and the packer must refuse to reorder such programs
# ---------------------------------------------------------------------------------------------


def test_faithful_mode_reproduces_int32_wraparound():
    from tsim_b200 import pack as PK

    rng = np.random.default_rng(77)
    F, n_c = 6, 2
    levels = []
    for k in range(n_c + 1):
        lv = random_level(rng, G=3, P=F + k, A=44, H=2, C=2, D=3, approx=False, density=0.4, power2_range=(-3, 3))
        lv.node_phases.counts[:] = 44
        lv.node_phases.phases[:] = rng.choice([1, 3, 5, 7], size=lv.node_phases.phases.shape)  # no vanishing factors
        lv.prefactor.floatfactor[:] = rng.integers(-9, 10, size=(3, 4))
        levels.append(lv)
    comp = CompiledComponent((0, 1), np.arange(F, dtype=np.int32), tuple(levels))
    prog = make_program([comp], num_f=F)
    ok, info = PK.reorder_is_exact(prog)
    assert not ok
    dp = _device_program(prog, "auto")
    assert dp.info["mode"] == 0
    with pytest.raises(ValueError):
        PK.pack_program(prog, mode="sliced")
    f = rng.integers(0, 2, size=(600, F)).astype(np.uint8)
    want, want_dev = _oracle(prog, f, (6, 6))
    got, dev = dp.sample(f, (6, 6))
    assert np.array_equal(got, want) and _dev_equal(dev, want_dev)
    # the wrap really happens: amplitudes differ from an int64 evaluation of the same node products
    x = np.hstack([f[:50], np.ones((50, 0), np.uint8)])
    amp = E.evaluate(levels[0], x)
    assert np.array_equal(dp.evaluate(0, 0, x).view(np.uint32), amp.view(np.uint32))


@pytest.mark.parametrize("mode", ["fast", "sliced"])
def test_reordered_modes_agree_with_faithful_near_the_bound(mode):
    from tsim_b200 import pack as PK

    rng = np.random.default_rng(5)
    F, n_c = 8, 2
    levels = []
    for k in range(n_c + 1):
        lv = random_level(rng, G=6, P=F + k, A=18, H=3, C=3, D=2, approx=False, density=0.35, power2_range=(-6, 0))
        lv.prefactor.floatfactor[:] = rng.integers(-3, 4, size=(6, 4))
        levels.append(lv)
    comp = CompiledComponent((0, 1), np.arange(F, dtype=np.int32), tuple(levels))
    prog = make_program([comp], num_f=F)
    ok, info = PK.reorder_is_exact(prog)
    assert ok and info["log2_bound"] > 22
    f = rng.integers(0, 2, size=(2000, F)).astype(np.uint8)
    want, want_dev = _oracle(prog, f, (2, 9))
    got, dev = _device_program(prog, mode).sample(f, (2, 9))
    assert np.array_equal(got, want) and _dev_equal(dev, want_dev)


@pytest.mark.parametrize("host_pack", ["0", "1"])
@pytest.mark.parametrize("name,B", [("cfg2_distill35", 600_001), ("cfg5_distill85", 70_000), ("cfg3_surface_d5", 300_000)])
def test_byte_rows_host_packed_or_dma(name, B, host_pack, monkeypatch):
    """The reference's uint8[B, num_f] rows reach the device either packed by host threads (1/8 of the PCIe bytes) or
    as bytes packed by K0; same bits either way (several pipeline slices, ragged tail, 1..3 words per row)."""
    monkeypatch.setenv("TSIM_B200_HOST_PACK", host_pack)
    prog = synthetic_program(name)
    nf = prog.infer_num_f()
    f = ChannelSampler.from_bit_probs(noise_probs(nf, 5e-3), seed=9).sample(B)
    f[-1, :] = 1
    dp = _device_program(prog, "auto")
    got, dev = dp.sample(f, (2, 3))
    packed, dev2 = dp.sample(pack_f_rows(f), (2, 3), packed_out=True)
    assert np.array_equal(np.packbits(got, axis=1, bitorder="little"), packed.view(np.uint8)[:, : (prog.num_outputs + 7) // 8])
    want = oracle.sample_program(prog, f[-1024:], (2, 3), shot_offset=B - 1024, check_norm=False)
    assert np.array_equal(got[-1024:], want)
