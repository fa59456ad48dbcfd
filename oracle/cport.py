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
        lib.tso_census.restype = C.c_int
        lib.tso_census.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def max_threads() -> int:
    return os.cpu_count() or 1


def _faithful_blob(program):
    from tsim_b200.pack import pack_program  # the blob format is shared with the product; the arithmetic is not

    packed = getattr(program, "_oracle_blob", None)
    if packed is None:
        packed = pack_program(program, mode="faithful")
        try:
            program._oracle_blob = packed
        except Exception:
            pass
    return packed


CENSUS_TOL = 2.0**-20
CENSUS_TOL_WIDE = 2.0**-12


def census(program, f_params: np.ndarray, key, *, shot_offset: int = 0, tol: float = CENSUS_TOL, tol_wide: float = CENSUS_TOL_WIDE, threads: int | None = None) -> dict:
    """Margin census of every Bernoulli draw of a batch (``tso_census`` in oracle/c/oracle.c).

    The oracle's float32 path decides the bits; each draw ``u < p1/prev`` is re-judged (a) with a float64 evaluation of
    the same amplitudes from the exact Z[w] values ("truth" for the stored inputs) and (b) with an alternative float32
    lowering (pairwise reduction over graphs, FMA-contracted products, ``sqrt(re^2+im^2)``).  A draw is a *margin draw*
    when ``|u - q| <= tol * q`` (``tol_wide``: a second, wider band counted in the same pass): only those can depend on how XLA orders / contracts the float32 tail
    (``evaluate.py:56-59``, ``abs(complex64)``), provided ``outside_margin_flips_f64`` is 0."""
    packed = _faithful_blob(program)
    f = np.ascontiguousarray(np.asarray(f_params).astype(np.uint8, copy=False))
    B = f.shape[0]
    blob = np.ascontiguousarray(packed.blob)
    lib = load()
    nf = f.shape[1]
    threads = max(1, int(threads or max_threads()))
    step = max(64, -(-B // (threads * 8))) if B else 1
    spans = [(lo, min(B, lo + step)) for lo in range(0, B, step)]

    def run(span):
        lo, hi = span
        st = np.zeros(8, dtype=np.int64)
        mr = np.zeros(2, dtype=np.float64)
        rc = lib.tso_census(blob.ctypes.data_as(C.c_void_p), C.c_void_p(f.ctypes.data + lo * nf), hi - lo, int(shot_offset) + lo,
                            int(key[0]), int(key[1]), float(tol), float(tol_wide), st.ctypes.data_as(C.c_void_p), mr.ctypes.data_as(C.c_void_p))
        if rc:
            raise RuntimeError(f"C oracle census rejected the program blob ({rc})")
        return st, mr

    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(run, spans))
    st = sum((p[0] for p in parts), np.zeros(8, dtype=np.int64))
    mr = np.max([p[1] for p in parts], axis=0) if parts else np.zeros(2)
    return {
        "shots": int(B),
        "draws": int(st[0]),
        "tol": float(tol),
        "margin_draws": int(st[1]),
        "outside_margin_flips_f64": int(st[2]),
        "flips_f64": int(st[3]),
        "flips_alt_f32_lowering": int(st[4]),
        "degenerate_draws": int(st[5]),
        "tol_wide": float(tol_wide),
        "margin_draws_wide": int(st[6]),
        "outside_wide_margin_flips_f64": int(st[7]),
        "max_rel_dev_f32_vs_f64": float(mr[0]),
        "max_rel_dev_alt_vs_f64": float(mr[1]),
    }


def sample_program(program, f_params: np.ndarray, key, *, shot_offset: int = 0, return_deviations: bool = False,
                   threads: int | None = None):
    """Same contract as ``oracle.sample_program(..., check_norm=False)``; ``program`` is a CompiledProgram.

    ``threads`` > 1 cuts the rows into slices that keep their in-batch RNG counters and runs them on a thread pool
    (the C call releases the GIL)."""
    packed = _faithful_blob(program)
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
