"""ctypes wrapper of the C restatement (oracle/c/oracle.c) -- test / baseline infrastructure only.

``-march=native`` ties the binary to the machine it was built on, so it is rebuilt (a second with gcc) whenever
the source is newer or the host CPU differs from the one recorded next to the binary."""

from __future__ import annotations

import ctypes as C
import os
import platform
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
STAMP = os.path.join(OUT_DIR, "host.txt")
_lib = None


def _host_id() -> str:
    cpu = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name") or line.startswith("flags"):
                    cpu += line
                    if line.startswith("flags"):
                        break
    except OSError:
        pass
    return platform.machine() + "\n" + cpu


def build(force: bool = False) -> str:
    fresh = (
        os.path.exists(LIB)
        and os.path.getmtime(LIB) >= os.path.getmtime(SRC)
        and os.path.exists(STAMP)
        and open(STAMP).read() == _host_id()
    )
    if fresh and not force:
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    res = subprocess.run(["make", "-B", "-C", os.path.join(HERE, "c")], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the C oracle failed:\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(_host_id())
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.tso_sample.restype = C.c_int
        lib.tso_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def max_threads() -> int:
    return os.cpu_count() or 1


def sample_program(program, f_params: np.ndarray, key, *, shot_offset: int = 0, return_deviations: bool = False,
                   threads: int | None = None):
    """Same contract as ``oracle.sample_program(..., check_norm=False)``; ``program`` is a CompiledProgram.

    ``threads`` > 1 cuts the rows into slices that keep their in-batch RNG counters and runs them on a thread pool
    (the C call releases the GIL)."""
    from tsim_b200.pack import pack_program  # the blob format is shared with the product; the arithmetic is not

    packed = getattr(program, "_oracle_blob", None)
    if packed is None:
        packed = pack_program(program, mode="faithful")
        try:
            program._oracle_blob = packed
        except Exception:
            pass
    f = np.ascontiguousarray(np.asarray(f_params).astype(np.uint8, copy=False))
    B = f.shape[0]
    out = np.zeros((B, packed.n_out), dtype=np.uint8)
    dev = np.zeros(max(1, packed.n_components), dtype=np.float32)
    blob = np.ascontiguousarray(packed.blob)
    lib = load()
    nf, no = f.shape[1], packed.n_out

    def run(lo: int, hi: int) -> int:
        return lib.tso_sample(
            blob.ctypes.data_as(C.c_void_p), C.c_void_p(f.ctypes.data + lo * nf), hi - lo, int(shot_offset) + lo,
            int(key[0]), int(key[1]), C.c_void_p(out.ctypes.data + lo * no), dev.ctypes.data_as(C.c_void_p),
        )

    threads = max(1, int(threads or 1))
    if threads == 1 or B < 2 * threads:
        rcs = [run(0, B)]
    else:
        from concurrent.futures import ThreadPoolExecutor

        step = max(64, -(-B // (threads * 8)))
        spans = [(lo, min(B, lo + step)) for lo in range(0, B, step)]
        with ThreadPoolExecutor(threads) as ex:
            rcs = list(ex.map(lambda sp: run(*sp), spans))
    if any(rcs):
        raise RuntimeError("C oracle rejected the program blob")
    bits = out.view(np.bool_)
    return (bits, list(dev[: packed.n_components])) if return_deviations else bits
