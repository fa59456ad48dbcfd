"""The C-ABI library builds, loads and exports every symbol declared in include/tsim_b200.h (no GPU needed)."""

import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tsim_b200.build import build
    from tsim_b200 import _lib

    build()
    return _lib.load()


def test_exports_match_header(lib):
    from tsim_b200 import _lib

    header = open(os.path.join(ROOT, "include", "tsim_b200.h")).read()
    declared = set(re.findall(r"\b(tsb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_split_key_matches_oracle(lib):
    import oracle
    from tsim_b200.backend import split_key

    key = (0, 0)
    for _ in range(5):
        assert split_key(key) == oracle.split(key)
        key = split_key(key)[0]
    assert split_key((0xDEADBEEF, 0x12345678)) == oracle.split((0xDEADBEEF, 0x12345678))


def test_no_device_fails_loudly(lib):
    import ctypes as C

    if lib.tsb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from tsim_b200.backend import DeviceProgram
    import kat_programs as K

    with pytest.raises(RuntimeError):
        DeviceProgram(K.hm_program())
    assert lib.tsb_host_alloc(16) is None


def test_host_side_bit_packers_agree_with_numpy():
    """The SIMD row packers of the end-to-end input path (csrc/host_pack.cpp; CPU code, no device needed): every ISA
    variant this host supports against np.packbits, ragged row widths and the over-read guard at the end of the array."""
    import ctypes as C

    import numpy as np

    from tsim_b200 import _lib
    from tsim_b200.noise import pack_f_rows

    lib = C.CDLL(_lib.LIB_PATH)
    lib.tsb_host_pack_rows.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_int]
    best = lib.tsb_host_pack_isa()
    rng = np.random.default_rng(0)
    for nf in (1, 7, 63, 64, 65, 121, 128, 160, 200):
        n = 777
        f = (rng.random((n, nf)) < 0.3).astype(np.uint8)
        f[rng.integers(0, n, 50), rng.integers(0, nf, 50)] = 255  # only the low bit counts, like the device's K0
        wf = max(1, (nf + 63) // 64)
        want = pack_f_rows(f & 1)
        for isa in range(best + 1):
            out = np.zeros((n, wf), np.uint64)
            lib.tsb_host_pack_rows(f.ctypes.data, n, nf, wf, out.ctypes.data, f.size, isa)
            assert np.array_equal(out, want), (nf, isa)


def test_layout_row_bytes_is_pure_host_arithmetic():
    # tsb_layout_row_bytes needs no device: sizes of the result arrays of tsb_sample_noisy_host_layout
    import ctypes as C

    from tsim_b200 import _lib

    lib = _lib.load()

    def size(segments, bit_packed, split=0):
        lay = _lib.TsbLayout.make(segments, bit_packed=bit_packed, split=split)
        return tuple(int(lib.tsb_layout_row_bytes(C.byref(lay), w)) for w in (0, 1))

    assert size([(0, 15)], True) == (2, 0)
    assert size([(15, 5), (0, 15), (15, 5)], True) == (4, 0)  # prepend + append observables: 25 bits
    assert size([(0, 15), (15, 5)], True, split=1) == (2, 1)  # separate_observables
    assert size([(0, 15), (15, 5)], False, split=1) == (15, 5)
    assert size([(0, 121)], True) == (16, 0)
    bad = _lib.TsbLayout.make([(0, 3)], bit_packed=True)
    bad.n_segments = 7
    assert int(lib.tsb_layout_row_bytes(C.byref(bad), 0)) == -1
