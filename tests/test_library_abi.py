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
