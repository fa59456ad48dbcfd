"""`install()` against a stand-in for the ``tsim.sampler`` module: the consumer below makes the same calls, in the same
order, as the reference's ``_sample_batches`` (src/tsim/sampler.py:388-414: ``jnp.asarray`` -> ``sample_program`` ->
``jnp.concatenate`` -> ``copy_d2h``), with a ``copy_d2h`` that, like the reference's on a CUDA box
(utils/cuda_helpers.py:120-141), dereferences ``unsafe_buffer_pointer()`` unless the source says it lives on the host.
CPU only: the device is the oracle-backed fake of test_sampler_host_logic."""

import sys
import types

import numpy as np
import pytest

import kat_programs as K
import oracle
import tsim_b200.sampler as S
from test_sampler_host_logic import FakeDeviceProgram


class _FakeJnp:
    uploads = 0

    @staticmethod
    def asarray(a):
        return np.asarray(a)

    @classmethod
    def concatenate(cls, arrays, axis=0):
        cls.uploads += 1  # a real jnp.concatenate would move host rows to the default device
        return np.concatenate(arrays, axis=axis)


def _strict_copy_d2h(src, *, dst=None):
    devices = getattr(src, "devices", None)
    on_host = devices is not None and any(getattr(d, "platform", "") == "cpu" for d in devices())
    if not on_host:
        src.unsafe_buffer_pointer()  # AttributeError for a plain ndarray, as in the reference on a GPU box
    out = np.empty(src.shape, dtype=src.dtype)
    out[:] = src
    return out


@pytest.fixture
def fake_tsim(monkeypatch):
    pkg, mod = types.ModuleType("tsim"), types.ModuleType("tsim.sampler")
    mod.sample_program = lambda *a, **k: (_ for _ in ()).throw(AssertionError("reference path must not run"))
    mod.copy_d2h = _strict_copy_d2h
    mod.jnp = _FakeJnp
    pkg.sampler = mod
    monkeypatch.setitem(sys.modules, "tsim", pkg)
    monkeypatch.setitem(sys.modules, "tsim.sampler", mod)
    monkeypatch.setattr(S, "DeviceProgram", FakeDeviceProgram)
    S._device_cache.clear()
    _FakeJnp.uploads = 0
    yield mod
    S.uninstall()
    S._device_cache.clear()


def _reference_style_batches(ts, program, f_batches, keys):
    """The call pattern of the reference's ``_sample_batches`` (names looked up on the module at call time)."""
    batches = []
    for f_np, key in zip(f_batches, keys):
        f_params = ts.jnp.asarray(f_np)
        batches.append(ts.sample_program(program, f_params, key))
    combined = batches[0] if len(batches) == 1 else ts.jnp.concatenate(batches, axis=0)
    return ts.copy_d2h(combined)


@pytest.mark.parametrize("n_batches", [1, 3])
def test_install_routes_host_rows_through_the_reference_call_pattern(fake_tsim, n_batches):
    prog = K.three_coin_program()
    assert S.install() and S.install()  # idempotent
    rng = np.random.default_rng(0)
    fs = [(rng.random((50, 1)) < 0.3).astype(np.uint8) for _ in range(n_batches)]
    keys = [(0, 10 + i) for i in range(n_batches)]
    out = _reference_style_batches(fake_tsim, prog, fs, keys)
    want = np.concatenate([oracle.sample_program(prog, f, k) for f, k in zip(fs, keys)], axis=0)
    assert isinstance(out, np.ndarray) and out.dtype == np.bool_ and np.array_equal(out, want)
    assert _FakeJnp.uploads == 0  # host rows were concatenated on the host
    S.uninstall()
    assert fake_tsim.copy_d2h is _strict_copy_d2h and fake_tsim.jnp is _FakeJnp


def test_result_passes_an_unpatched_copy_d2h(fake_tsim):
    """Even without the rebinding, the returned rows declare themselves host-resident."""
    S.install()
    prog = K.three_coin_program()
    bits = fake_tsim.sample_program(prog, np.zeros((10, 1), np.uint8), (0, 1))
    assert np.array_equal(_strict_copy_d2h(bits), oracle.sample_program(prog, np.zeros((10, 1), np.uint8), (0, 1)))


def test_wider_f_rows_than_the_program_references(fake_tsim):
    # the channel sampler's signature matrix may have trailing columns no output depends on: ignored, as in the reference
    S.install()
    prog = K.three_coin_program()
    assert prog.infer_num_f() == 1
    f = np.zeros((20, 4), np.uint8)
    f[:, 2] = 1
    bits = fake_tsim.sample_program(prog, f, (0, 3))
    assert np.array_equal(bits, oracle.sample_program(prog, f, (0, 3)))
