"""ctypes binding of ``libtsim_b200.so`` (C ABI: ``include/tsim_b200.h``).

There is no CPU implementation behind this module: if the library is missing the import of the
product path fails, and without a CUDA device every compute call raises ``RuntimeError``.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSIM_B200_LIB") or os.path.join(_HERE, "libtsim_b200.so")  # override: kernel experiments

TSB_F_BYTES, TSB_F_PACKED = 0, 1
TSB_OUT_BYTES, TSB_OUT_PACKED = 0, 1

EXPORTS = (
    "tsb_last_error",
    "tsb_device_count",
    "tsb_program_create",
    "tsb_program_destroy",
    "tsb_program_info",
    "tsb_split_key",
    "tsb_sample_device",
    "tsb_sample_host",
    "tsb_evaluate_host",
    "tsb_pack_f_device",
    "tsb_unpack_out_device",
    "tsb_last_kernel_ms",
    "tsb_host_alloc",
    "tsb_host_free",
    "tsb_noise_create",
    "tsb_noise_destroy",
    "tsb_noise_sample_device",
    "tsb_noise_sample_host",
    "tsb_sample_noisy_host",
    "tsb_sample_noisy_host_layout",
    "tsb_layout_row_bytes",
    "tsb_device_mem_info",
    "tsb_program_set_pattern_cache",
    "tsb_program_set_aux",
    "tsb_postselect_create",
    "tsb_postselect_push_host",
    "tsb_postselect_push_noise",
    "tsb_postselect_dispatch",
    "tsb_postselect_finish",
    "tsb_postselect_destroy",
    "tsb_memcpy_peer_async",
)


class TsbInfo(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("words", C.c_int32),
        ("num_f", C.c_int32),
        ("num_outputs", C.c_int32),
        ("n_direct", C.c_int32),
        ("n_components", C.c_int32),
        ("n_draws", C.c_int32),
        ("words_f64", C.c_int32),
        ("words_out64", C.c_int32),
        ("resident", C.c_int32),
        ("n_chunks", C.c_int32),
        ("smem_bytes", C.c_int32),
        ("threads", C.c_int32),
        ("grid", C.c_int32),
        ("data_bytes", C.c_int64),
    ]

    def as_dict(self) -> dict:
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class TsbLayout(C.Structure):
    """``tsb_layout``: column ranges of the result, concatenated along the bit axis (include/tsim_b200.h)."""

    _fields_ = [
        ("n_segments", C.c_int32),
        ("lo", C.c_int32 * 4),
        ("n", C.c_int32 * 4),
        ("bit_packed", C.c_int32),
        ("split", C.c_int32),
    ]

    @classmethod
    def make(cls, segments, *, bit_packed: bool, split: int = 0) -> "TsbLayout":
        segments = [(int(lo), int(n)) for lo, n in segments]
        if not 0 <= len(segments) <= 4:
            raise ValueError("a layout has at most four column ranges")
        out = cls()
        out.n_segments = len(segments)
        for i, (lo, n) in enumerate(segments):
            out.lo[i], out.n[i] = lo, n
        if not 0 <= int(split) <= len(segments):
            raise ValueError("split must be a number of leading column ranges")
        out.bit_packed, out.split = int(bool(bit_packed)), int(split)
        return out


_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m tsim_b200.build` (nvcc, sm_100a). "
            "tsim_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, u32, i64, i32 = C.c_void_p, C.c_uint32, C.c_int64, C.c_int
    lib.tsb_last_error.restype = C.c_char_p
    lib.tsb_last_error.argtypes = []
    lib.tsb_device_count.restype = i32
    lib.tsb_device_count.argtypes = []
    lib.tsb_program_create.restype = i32
    lib.tsb_program_create.argtypes = [vp, C.c_size_t, i32, C.POINTER(vp)]
    lib.tsb_program_destroy.restype = i32
    lib.tsb_program_destroy.argtypes = [vp]
    lib.tsb_program_info.restype = i32
    lib.tsb_program_info.argtypes = [vp, C.POINTER(TsbInfo)]
    lib.tsb_split_key.restype = None
    lib.tsb_split_key.argtypes = [u32, u32, C.POINTER(u32 * 4)]
    lib.tsb_sample_device.restype = i32
    lib.tsb_sample_device.argtypes = [vp, vp, i64, i64, u32, u32, vp, vp, vp]
    lib.tsb_sample_host.restype = i32
    lib.tsb_sample_host.argtypes = [vp, vp, i32, i64, i64, u32, u32, vp, i32, vp]
    lib.tsb_evaluate_host.restype = i32
    lib.tsb_evaluate_host.argtypes = [vp, i32, i32, vp, i64, vp]
    lib.tsb_pack_f_device.restype = i32
    lib.tsb_pack_f_device.argtypes = [vp, vp, i64, vp, vp]
    lib.tsb_unpack_out_device.restype = i32
    lib.tsb_unpack_out_device.argtypes = [vp, vp, i64, vp, vp]
    lib.tsb_last_kernel_ms.restype = C.c_float
    lib.tsb_last_kernel_ms.argtypes = [vp, C.POINTER(i32)]
    lib.tsb_host_alloc.restype = vp
    lib.tsb_host_alloc.argtypes = [C.c_size_t]
    lib.tsb_host_free.restype = None
    lib.tsb_host_free.argtypes = [vp]
    u64 = C.c_uint64
    lib.tsb_noise_create.restype = i32
    lib.tsb_noise_create.argtypes = [i32, vp, vp, vp, i32, i32, C.POINTER(vp)]
    lib.tsb_noise_destroy.restype = i32
    lib.tsb_noise_destroy.argtypes = [vp]
    lib.tsb_noise_sample_device.restype = i32
    lib.tsb_noise_sample_device.argtypes = [vp, i64, i64, u64, u64, i32, vp, vp]
    lib.tsb_noise_sample_host.restype = i32
    lib.tsb_noise_sample_host.argtypes = [vp, i64, i64, u64, u64, i32, vp]
    lib.tsb_sample_noisy_host.restype = i32
    lib.tsb_sample_noisy_host.argtypes = [vp, vp, i64, i64, u32, u32, u64, u64, i32, vp, i32, vp, vp]
    lib.tsb_sample_noisy_host_layout.restype = i32
    lib.tsb_sample_noisy_host_layout.argtypes = [vp, vp, i64, i64, u32, u32, u64, u64, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.tsb_device_mem_info.restype = i32
    lib.tsb_device_mem_info.argtypes = [i32, C.POINTER(i64), C.POINTER(i64)]
    lib.tsb_layout_row_bytes.restype = i64
    lib.tsb_layout_row_bytes.argtypes = [vp, i32]
    lib.tsb_program_set_aux.restype = i32
    lib.tsb_program_set_aux.argtypes = [vp, vp]
    lib.tsb_program_set_pattern_cache.restype = i32
    lib.tsb_program_set_pattern_cache.argtypes = [vp, i32, i64, C.POINTER(i64)]
    lib.tsb_postselect_create.restype = i32
    lib.tsb_postselect_create.argtypes = [vp, i64, i64, vp, vp, i32, C.POINTER(vp)]
    lib.tsb_postselect_push_host.restype = i32
    lib.tsb_postselect_push_host.argtypes = [vp, vp, i64, C.POINTER(i64)]
    lib.tsb_postselect_push_noise.restype = i32
    lib.tsb_postselect_push_noise.argtypes = [vp, vp, i64, u64, u64, C.POINTER(i64)]
    lib.tsb_postselect_dispatch.restype = i32
    lib.tsb_postselect_dispatch.argtypes = [vp, u32, u32, i32, vp, C.POINTER(i64)]
    lib.tsb_postselect_finish.restype = i32
    lib.tsb_postselect_finish.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.tsb_postselect_destroy.restype = i32
    lib.tsb_postselect_destroy.argtypes = [vp]
    lib.tsb_memcpy_peer_async.restype = i32
    lib.tsb_memcpy_peer_async.argtypes = [vp, vp, C.c_size_t, vp]
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Map a tsb_status to the exception classes the reference raises (cuda_helpers.py:61,140)."""
    if rc == 0:
        return
    msg = load().tsb_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(f"tsim_b200: {msg}")
    if rc == -3:
        raise NotImplementedError(f"tsim_b200: {msg}")
    raise RuntimeError(f"tsim_b200: {msg}")
