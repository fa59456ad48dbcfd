"""Device-resident program handle: upload once, sample / evaluate many times."""

from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Any

import numpy as np

from . import _lib
from .pack import MODE_SLICED, PackedProgram, pack_program
from .program import CompiledProgram, from_tsim


PATTERN_CACHE_AUTO_WEIGHT = 3  # "auto": up to three flipped selected f bits per component (table capped by the library)


def key_words(key: Any) -> tuple[int, int]:
    """Accept ``(k0, k1)``, a uint32[2] array or a jax PRNG key (typed or raw)."""
    if isinstance(key, tuple) and len(key) == 2:
        return int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    arr = None
    try:
        arr = np.asarray(key)
        if arr.dtype == object or arr.shape != (2,):
            arr = None
    except Exception:
        arr = None
    if arr is None:
        import jax  # only reached for typed jax keys

        arr = np.asarray(jax.random.key_data(key))
    arr = arr.astype(np.uint32).reshape(2)
    return int(arr[0]), int(arr[1])


def split_key(key: tuple[int, int]) -> tuple[tuple[int, int], tuple[int, int]]:
    """``carry, sub = jax.random.split(key)`` (threefry2x32, host side of the library)."""
    out = (C.c_uint32 * 4)()
    _lib.load().tsb_split_key(key[0], key[1], C.byref(out))
    return (int(out[0]), int(out[1])), (int(out[2]), int(out[3]))


class PinnedArray:
    """Page-locked host buffer exposed as a NumPy array (reference: ``alloc_pinned_numpy``)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        lib = _lib.load()
        self._ptr = lib.tsb_host_alloc(max(nbytes, 1))
        if not self._ptr:
            raise RuntimeError("tsim_b200: " + lib.tsb_last_error().decode())
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=nbytes).view(self.dtype).reshape(self.shape)
        self._fin = weakref.finalize(self, lib.tsb_host_free, self._ptr)


class PinnedPool:
    """Recycles page-locked host buffers for results.

    The reference allocates a fresh ``cudaHostAlloc`` region for every result and frees it when the
    returned array dies (``cuda_helpers.py:48-102``).  Page-locking is slow (milliseconds), so here a
    dead result's region goes back to a free list instead; an array and every view derived from it keep
    the region checked out through their common ctypes base object.
    """

    GRANULE = 1 << 20

    def __init__(self):
        self._free: dict[int, list[int]] = {}
        self._lib = None

    def _give_back(self, cap: int, ptr: int) -> None:
        self._free.setdefault(cap, []).append(ptr)

    def take(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        cap = max(self.GRANULE, -(-nbytes // self.GRANULE) * self.GRANULE)
        if self._lib is None:
            self._lib = _lib.load()
        stack = self._free.get(cap)
        if stack:
            ptr = stack.pop()
        else:
            ptr = self._lib.tsb_host_alloc(cap)
            if not ptr:
                raise RuntimeError("tsim_b200: " + self._lib.tsb_last_error().decode())
        carr = (C.c_uint8 * cap).from_address(ptr)
        weakref.finalize(carr, self._give_back, cap, ptr)
        return np.frombuffer(carr, dtype=np.uint8, count=nbytes).view(dtype).reshape(shape)


_result_pool = PinnedPool()


class DeviceProgram:
    """A compiled program uploaded to one GPU (``tsb_program`` handle)."""

    def __init__(self, program: CompiledProgram | PackedProgram | Any, *, device: int = 0, mode: str = "auto", joint: bool = False,
                 pattern_cache: int | None | str = "default"):
        """``pattern_cache``: ``"default"`` = the ``TSIM_B200_PATTERN_CACHE`` environment variable (``off`` or a weight 0..3),
        else ``"auto"`` for sliced programs (the largest weight whose table stays L2-sized) and off for per-row ones;
        an int or ``None`` sets it explicitly.  The cache never changes sampled bits (see ``set_pattern_cache``)."""
        lib = _lib.load()
        self._lib = lib
        self.device = int(device)
        self.pattern_cache = None
        if isinstance(mode, str) and mode.endswith("-direct"):  # e.g. "sliced-direct": the records of `mode`, no pattern cache
            mode = mode[: -len("-direct")]
            pattern_cache = None
        self._aux = None
        if isinstance(program, PackedProgram):
            packed = program
            self.program = None
        else:
            self.program = from_tsim(program)
            packed = pack_program(self.program, mode=mode, joint=joint)
        self.packed = packed
        self.joint = bool(packed.stats.get("joint", joint))
        self._h, self._fin = self._create(packed)
        if packed.mode == MODE_SLICED:
            if self.program is None:
                raise ValueError("a sliced program needs the CompiledProgram to build its per-row companion")
            # companion per-row program: normalisation check of shot 0, evaluate(), pattern cache
            self._aux = DeviceProgram(self.program, device=device, mode="rowwise")
            _lib.check(lib.tsb_program_set_aux(self._h, self._aux._h))
        info = _lib.TsbInfo()
        _lib.check(lib.tsb_program_info(self._h, C.byref(info)))
        self.info = info.as_dict()
        if pattern_cache == "default":
            env = os.environ.get("TSIM_B200_PATTERN_CACHE", "").strip().lower()
            if env in ("off", "none", "-1"):
                pattern_cache = None
            elif env.isdigit():
                pattern_cache = int(env)
            else:
                pattern_cache = "auto" if packed.mode == MODE_SLICED else None
        if self.joint or self.info["n_draws"] == 0:
            pattern_cache = None
        if pattern_cache == "auto":
            pattern_cache = PATTERN_CACHE_AUTO_WEIGHT
        if pattern_cache is not None:
            self.set_pattern_cache(int(pattern_cache))

    def _create(self, packed: PackedProgram):
        blob = np.ascontiguousarray(packed.blob, dtype=np.uint32)
        handle = C.c_void_p()
        _lib.check(self._lib.tsb_program_create(blob.ctypes.data_as(C.c_void_p), blob.size, self.device, C.byref(handle)))
        return handle, weakref.finalize(self, self._lib.tsb_program_destroy, handle)

    # ---------------------------------------------------------------------------------------
    @property
    def num_f(self) -> int:
        return self.info["num_f"]

    @property
    def num_outputs(self) -> int:
        return self.info["num_outputs"]

    def set_pattern_cache(self, max_weight: int | None, max_entries: int = 0) -> int:
        """Tabulate the probability trees ``|E_k(pattern, prefix)|`` of all selected-f patterns of weight <= ``max_weight``
        (0..3; per component the largest weight <= ``max_weight`` whose table fits ``max_entries``; ``None`` switches the
        cache off).  The chain rule of ``_sample_component`` (reference ``sampler.py:45-81``) is a pure function of the
        selected f bits, so light shots walk the table and draw; the rest take the full evaluation (for a sliced program
        K0t / K1s / K2a over the list of remaining rows).  Purely a speed-up: sampled bits do not change.  Returns the
        table size in entries."""
        n = C.c_int64(0)
        _lib.check(self._lib.tsb_program_set_pattern_cache(self._h, -1 if max_weight is None else int(max_weight), int(max_entries), C.byref(n)))
        self.pattern_cache = max_weight
        return int(n.value)

    def _active(self):
        """Handle that samples."""
        return self._h

    def close(self) -> None:
        self._fin()

    def last_kernel_ms(self) -> tuple[float, int]:
        n = C.c_int(0)
        ms = self._lib.tsb_last_kernel_ms(self._active(), C.byref(n))
        return float(ms), int(n.value)

    # ---------------------------------------------------------------------------------------
    def sample(self, f_params: np.ndarray, key, *, shot_offset: int = 0, packed_out: bool = False, out: np.ndarray | None = None):
        """One batch through the device.

        ``f_params``: ``uint8/bool[B, num_f]`` (reference format) or ``uint64[B, ceil(num_f/64)]`` packed.
        Returns ``(bits, norm_dev)``: ``bool[B, num_outputs]`` (or packed ``uint64`` rows) and the
        per-component maximum normalisation deviation of shot 0 (``float32[n_components]``).
        """
        if self.joint:
            raise ValueError("a joint-mode program can only be evaluated, not sampled")
        f = np.asarray(f_params)
        if f.ndim != 2:
            raise ValueError("f_params must be 2-D")
        B = f.shape[0]
        if f.dtype == np.uint64:
            if f.shape[1] != self.info["words_f64"]:
                raise ValueError(f"packed f_params must have {self.info['words_f64']} words per row")
            fmt = _lib.TSB_F_PACKED
        else:
            if f.shape[1] != self.num_f:
                raise ValueError(f"f_params must have {self.num_f} columns, got {f.shape[1]}")
            if f.dtype != np.uint8 and f.dtype != np.bool_:
                f = f.astype(np.uint8)
            fmt = _lib.TSB_F_BYTES
        f = np.ascontiguousarray(f)
        k0, k1 = key_words(key)
        if packed_out:
            shape, dtype, ofmt = (B, self.info["words_out64"]), np.uint64, _lib.TSB_OUT_PACKED
        else:
            shape, dtype, ofmt = (B, self.num_outputs), np.bool_, _lib.TSB_OUT_BYTES
        if out is None:
            out = _result_pool.take(shape, dtype) if B > 0 else np.empty(shape, dtype=dtype)
        elif out.shape != shape or out.dtype != dtype or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape, dtype or layout")
        dev = np.zeros(max(1, self.info["n_components"]), dtype=np.float32)
        _lib.check(
            self._lib.tsb_sample_host(
                self._active(),
                f.ctypes.data_as(C.c_void_p),
                fmt,
                B,
                int(shot_offset),
                k0,
                k1,
                out.ctypes.data_as(C.c_void_p),
                ofmt,
                dev.ctypes.data_as(C.c_void_p),
            )
        )
        return out, dev[: self.info["n_components"]]

    def sample_noisy(self, noise, B: int, key, *, shot_offset: int = 0, call: int | None = None, skip_shot0: bool = False,
                     packed_out: bool = False, return_f: bool = False):
        """Noise sampling + program sampling in one device pipeline (``noise``: ``DeviceChannelSampler``).

        Returns ``(bits, norm_dev)`` or ``(bits, norm_dev, f_packed)`` with ``return_f``."""
        if self.joint:
            raise ValueError("a joint-mode program can only be evaluated, not sampled")
        if noise.num_f != self.num_f:
            raise ValueError(f"noise sampler produces {noise.num_f} f bits, the program expects {self.num_f}")
        if call is None:
            call = noise.next_call()
        k0, k1 = key_words(key)
        B = int(B)
        if packed_out:
            shape, dtype, ofmt = (B, self.info["words_out64"]), np.uint64, _lib.TSB_OUT_PACKED
        else:
            shape, dtype, ofmt = (B, self.num_outputs), np.bool_, _lib.TSB_OUT_BYTES
        out = _result_pool.take(shape, dtype) if B > 0 else np.empty(shape, dtype=dtype)
        f_out = np.zeros((B, self.info["words_f64"]), dtype=np.uint64) if return_f else None
        dev = np.zeros(max(1, self.info["n_components"]), dtype=np.float32)
        _lib.check(
            self._lib.tsb_sample_noisy_host(
                self._active(), noise._h, B, int(shot_offset), k0, k1, int(noise.seed), int(call), int(skip_shot0),
                out.ctypes.data_as(C.c_void_p), ofmt, dev.ctypes.data_as(C.c_void_p),
                f_out.ctypes.data_as(C.c_void_p) if return_f else None,
            )
        )
        dev = dev[: self.info["n_components"]]
        return (out, dev, f_out) if return_f else (out, dev)

    def mem_info(self) -> tuple[int, int]:
        """(free, total) bytes of this program's device."""
        free, total = C.c_int64(0), C.c_int64(0)
        _lib.check(self._lib.tsb_device_mem_info(int(self.device), C.byref(free), C.byref(total)))
        return int(free.value), int(total.value)

    def layout_row_bytes(self, segments, *, bit_packed: bool, split: int = 0) -> tuple[int, int]:
        """Bytes per row of the (first, second) result array of a column layout (``tsb_layout_row_bytes``)."""
        lay = _lib.TsbLayout.make(segments, bit_packed=bit_packed, split=split)
        n = tuple(int(self._lib.tsb_layout_row_bytes(C.byref(lay), w)) for w in (0, 1))
        if min(n) < 0:
            raise ValueError("bad column layout")
        return n

    def sample_noisy_layout(self, noise, B: int, key, segments, *, bit_packed: bool, split: int = 0,
                            xor_row: np.ndarray | None = None, ref_mask: np.ndarray | None = None, call: int | None = None,
                            skip_shot0: bool = False, out: np.ndarray | None = None, out2: np.ndarray | None = None):
        """``sample_noisy`` with the result already in the caller's column layout (``tsb_sample_noisy_host_layout``):
        ``segments`` = column ranges ``(lo, n)`` concatenated along the bit axis, as ``np.packbits(bitorder="little")``
        bytes (``bit_packed``) or one bool byte per column; ``split > 0`` returns the first ``split`` ranges and the rest
        as two arrays.  ``xor_row`` (``uint64[words_out64]``) is XORed into every row first; with ``ref_mask`` shot 0 is
        the reference sample: ``row0 & ref_mask`` is XORed in as well and row 0 is returned instead of being part of
        the result.  -> ``(uint8[rows, row_bytes], uint8[rows, row_bytes2] | None, norm_dev, row0 | None)``."""
        if self.joint:
            raise ValueError("a joint-mode program can only be evaluated, not sampled")
        if noise.num_f != self.num_f:
            raise ValueError(f"noise sampler produces {noise.num_f} f bits, the program expects {self.num_f}")
        if call is None:
            call = noise.next_call()
        k0, k1 = key_words(key)
        B = int(B)
        lay = _lib.TsbLayout.make(segments, bit_packed=bit_packed, split=split)
        rb = self.layout_row_bytes(segments, bit_packed=bit_packed, split=split)
        rows = max(0, B - (1 if ref_mask is not None else 0))
        two = 0 < split < len(segments)

        def result(buf, row_bytes):
            if buf is None:
                return _result_pool.take((rows, row_bytes), np.uint8) if rows * row_bytes > 0 else np.empty((rows, row_bytes), np.uint8)
            if buf.dtype != np.uint8 or buf.ndim != 2 or buf.shape[0] < rows or buf.shape[1] != row_bytes or not buf.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous uint8[>= rows, row_bytes] array")
            return buf

        out = result(out, rb[0])
        out2 = result(out2, rb[1]) if two else None
        wo = self.info["words_out64"]

        def row_arg(a):
            if a is None:
                return None, None
            a = np.ascontiguousarray(np.asarray(a, dtype=np.uint64).reshape(-1))
            if a.shape[0] != wo:
                raise ValueError(f"packed output rows have {wo} words")
            return a, a.ctypes.data_as(C.c_void_p)

        xr, xr_p = row_arg(xor_row)
        rm, rm_p = row_arg(ref_mask)
        row0 = np.zeros(wo, dtype=np.uint64) if ref_mask is not None else None
        dev = np.zeros(max(1, self.info["n_components"]), dtype=np.float32)
        _lib.check(
            self._lib.tsb_sample_noisy_host_layout(
                self._active(), noise._h, B, 0, k0, k1, int(noise.seed), int(call), int(skip_shot0), C.byref(lay), xr_p, rm_p,
                out.ctypes.data_as(C.c_void_p), out2.ctypes.data_as(C.c_void_p) if out2 is not None else None,
                row0.ctypes.data_as(C.c_void_p) if row0 is not None else None, dev.ctypes.data_as(C.c_void_p),
            )
        )
        return out[:rows], (out2[:rows] if out2 is not None else None), dev[: self.info["n_components"]], row0

    def sample_device(self, d_f: int, B: int, key, d_out: int, *, shot_offset: int = 0, d_norm_dev: int = 0, stream: int = 0) -> None:
        """Launch on device pointers (e.g. ``torch.Tensor.data_ptr()``); asynchronous on ``stream``."""
        k0, k1 = key_words(key)
        _lib.check(
            self._lib.tsb_sample_device(
                self._active(), C.c_void_p(d_f), int(B), int(shot_offset), k0, k1, C.c_void_p(d_out),
                C.c_void_p(d_norm_dev) if d_norm_dev else None, C.c_void_p(stream) if stream else None,
            )
        )

    def postselect_session(self, shots: int, batch_size: int, mask: np.ndarray, ref: np.ndarray | None, num_detectors: int) -> "PostselectSession":
        """Device-side survivor buffering for ``_sample_batches_with_postselection`` (reference sampler.py:422-545)."""
        return PostselectSession(self, shots, batch_size, mask, ref, num_detectors)

    def level_params(self, component: int, level: int) -> int:
        """Number of parameters of ``components[component].compiled_scalar_graphs[level]``."""
        from . import pack as PK

        b = self.packed.blob
        if not 0 <= component < self.info["n_components"]:
            raise ValueError("component out of range")
        comp = b[int(b[PK.H_OFF_COMP]) + component * PK.COMP_WORDS :]
        if not 0 <= level < int(comp[5]):
            raise ValueError("level out of range")
        return int(b[int(b[PK.H_OFF_LEVEL]) + (int(comp[4]) + level) * PK.LEVEL_WORDS + 1])

    def evaluate(self, component: int, level: int, params: np.ndarray) -> np.ndarray:
        """``evaluate(components[component].compiled_scalar_graphs[level], params)`` -> complex64[B]."""
        x = np.ascontiguousarray(np.asarray(params).astype(np.uint8))
        if x.ndim != 2:
            raise ValueError("params must be 2-D")
        want = self.level_params(component, level)
        if x.shape[1] != want:
            raise ValueError(f"params must have {want} columns, got {x.shape[1]}")
        amp = np.zeros((x.shape[0], 2), dtype=np.float32)
        _lib.check(
            self._lib.tsb_evaluate_host(
                self._h, int(component), int(level), x.ctypes.data_as(C.c_void_p), x.shape[0], amp.ctypes.data_as(C.c_void_p)
            )
        )
        return amp.view(np.complex64).reshape(-1)


def _pack_row(bits: np.ndarray, words: int) -> np.ndarray:
    """``bool[n]`` -> ``uint64[words]`` (bit j = column j)."""
    b = np.zeros(words * 64, dtype=np.uint8)
    b[: len(bits)] = np.asarray(bits, dtype=np.uint8)
    return np.packbits(b, bitorder="little").view(np.uint64).copy()


class PostselectSession:
    """``tsb_postselect``: direct-detector test, order-preserving survivor compaction, fixed-shape batches and the
    scatter of sampled rows all stay on the GPU; the caller keeps the reference's loop and key schedule."""

    def __init__(self, dp: DeviceProgram, shots: int, batch_size: int, mask: np.ndarray, ref: np.ndarray | None, num_detectors: int):
        self.dp, self.shots, self.batch_size = dp, int(shots), int(batch_size)
        self._lib = dp._lib
        wo = dp.info["words_out64"]
        n_out = dp.num_outputs
        m = np.zeros(n_out, dtype=np.bool_)
        m[: len(mask)] = mask
        self._mask = _pack_row(m, wo)
        self._ref = None
        if ref is not None:
            r = np.zeros(n_out, dtype=np.bool_)
            r[: len(ref)] = ref
            self._ref = _pack_row(r, wo)
        h = C.c_void_p()
        _lib.check(self._lib.tsb_postselect_create(
            dp._h, self.shots, self.batch_size, self._mask.ctypes.data_as(C.c_void_p),
            self._ref.ctypes.data_as(C.c_void_p) if self._ref is not None else None, int(num_detectors), C.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, self._lib.tsb_postselect_destroy, h)
        self.dispatches = 0

    def push_host(self, f: np.ndarray) -> int:
        """A chunk of f rows (``uint8[n, num_f]`` or packed ``uint64``) -> number of survivors pending."""
        f = np.asarray(f)
        if f.dtype != np.uint64:
            from .noise import pack_f_rows

            f = pack_f_rows(f)
        f = np.ascontiguousarray(f)
        pending = C.c_int64(0)
        _lib.check(self._lib.tsb_postselect_push_host(self._h, f.ctypes.data_as(C.c_void_p), f.shape[0], C.byref(pending)))
        return int(pending.value)

    def push_noise(self, noise, n: int) -> int:
        pending = C.c_int64(0)
        _lib.check(self._lib.tsb_postselect_push_noise(self._h, noise._h, int(n), int(noise.seed), int(noise.next_call()), C.byref(pending)))
        return int(pending.value)

    def dispatch(self, key, *, final: bool = False) -> tuple[int, np.ndarray]:
        k0, k1 = key_words(key)
        n_comp = self.dp.info["n_components"]
        dev = np.zeros(max(1, n_comp), dtype=np.float32)
        pending = C.c_int64(0)
        _lib.check(self._lib.tsb_postselect_dispatch(self._h, k0, k1, int(final), dev.ctypes.data_as(C.c_void_p), C.byref(pending)))
        self.dispatches += 1
        return int(pending.value), dev[:n_comp]

    def finish(self, xor_kept: np.ndarray | None = None, xor_discarded: np.ndarray | None = None) -> tuple[np.ndarray, np.ndarray]:
        """-> ``(bool[shots, n_out], bool[shots] was_discarded)``; the two rows are XORed into kept / discarded shots."""
        wo, n_out = self.dp.info["words_out64"], self.dp.num_outputs
        out = _result_pool.take((self.shots, n_out), np.bool_) if self.shots else np.empty((0, n_out), np.bool_)
        disc = np.zeros(self.shots, dtype=np.bool_)
        xk = xd = None
        if xor_kept is not None or xor_discarded is not None:
            xk = _pack_row(xor_kept if xor_kept is not None else np.zeros(n_out, bool), wo)
            xd = _pack_row(xor_discarded if xor_discarded is not None else np.zeros(n_out, bool), wo)
        _lib.check(self._lib.tsb_postselect_finish(
            self._h, xk.ctypes.data_as(C.c_void_p) if xk is not None else None, xd.ctypes.data_as(C.c_void_p) if xd is not None else None,
            out.ctypes.data_as(C.c_void_p), _lib.TSB_OUT_BYTES, disc.ctypes.data_as(C.c_void_p)))
        self._fin()
        return out, disc


class MultiDeviceProgram:
    """The same program on several GPUs of one box, driven from ONE process (reference: single process, ``jax.devices()[0]``,
    sampler.py:310; SURVEY section 5 ``devices=``).  A batch is cut into contiguous row ranges (``shard.shard_range``), every
    device samples its range with ``shot_offset = lo`` -- the RNG counter is the in-batch shot index, so the bits do not
    depend on the number of devices -- and writes into its slice of one pinned host buffer.  The per-device calls run on a
    thread pool (ctypes releases the GIL; each handle owns its streams).  Same surface as :class:`DeviceProgram`."""

    def __init__(self, program, *, devices, mode: str = "auto", joint: bool = False, pattern_cache: int | None | str = "default"):
        from concurrent.futures import ThreadPoolExecutor

        devices = [int(d) for d in devices]
        if not devices:
            raise ValueError("devices must name at least one GPU")
        n_dev = _lib.load().tsb_device_count()
        for d in devices:
            if not 0 <= d < n_dev:
                raise ValueError(f"no such CUDA device: {d} (the box has {n_dev})")
        self.devices = devices
        self.parts = [DeviceProgram(program, device=d, mode=mode, joint=joint, pattern_cache=pattern_cache) for d in devices]
        first = self.parts[0]
        self.program, self.packed, self.joint, self.info, self.device = first.program, first.packed, first.joint, first.info, first.device
        self._pool = ThreadPoolExecutor(len(devices))

    num_f = property(lambda self: self.info["num_f"])
    num_outputs = property(lambda self: self.info["num_outputs"])
    pattern_cache = property(lambda self: self.parts[0].pattern_cache)

    def set_pattern_cache(self, max_weight, max_entries: int = 0) -> int:
        return [p.set_pattern_cache(max_weight, max_entries) for p in self.parts][0]

    def last_kernel_ms(self):
        return self.parts[0].last_kernel_ms()

    def _ranges(self, B: int):
        from .shard import shard_range

        return [shard_range(B, r, len(self.parts)) for r in range(len(self.parts))]

    def _result(self, B: int, packed_out: bool, out):
        if packed_out:
            shape, dtype = (B, self.info["words_out64"]), np.uint64
        else:
            shape, dtype = (B, self.num_outputs), np.bool_
        if out is None:
            out = _result_pool.take(shape, dtype) if B > 0 else np.empty(shape, dtype=dtype)
        elif out.shape != shape or out.dtype != dtype or not out.flags.c_contiguous:
            raise ValueError("out has the wrong shape, dtype or layout")
        return out

    def _gather(self, futures, shot_offset: int):
        devs = np.zeros(self.info["n_components"], dtype=np.float32)
        for (lo, hi), fut in futures:
            res = fut.result()
            if hi > lo and lo + shot_offset == 0:  # the norm check belongs to in-batch shot 0 (sampler.py:66-72)
                devs = np.asarray(res[1], dtype=np.float32)
        return devs

    def sample(self, f_params: np.ndarray, key, *, shot_offset: int = 0, packed_out: bool = False, out: np.ndarray | None = None):
        f = np.asarray(f_params)
        if f.ndim != 2:
            raise ValueError("f_params must be 2-D")
        B = f.shape[0]
        out = self._result(B, packed_out, out)
        futures = []
        for part, (lo, hi) in zip(self.parts, self._ranges(B)):
            if hi > lo:
                futures.append(((lo, hi), self._pool.submit(part.sample, f[lo:hi], key, shot_offset=shot_offset + lo, packed_out=packed_out, out=out[lo:hi])))
        return out, self._gather(futures, shot_offset)

    def sample_noisy(self, noise, B: int, key, *, shot_offset: int = 0, call: int | None = None, skip_shot0: bool = False,
                     packed_out: bool = False, return_f: bool = False):
        """``noise``: a :class:`tsim_b200.noise.MultiDeviceChannelSampler` over the same devices."""
        if return_f:
            raise ValueError("return_f is a single-device debugging aid")
        B = int(B)
        if call is None:
            call = noise.next_call()
        out = self._result(B, packed_out, None)

        def run(part, nz, lo, hi):
            # a shard's rows land in its own pinned result; copy into the common buffer (8 B per shot when packed)
            res = part.sample_noisy(nz, hi - lo, key, shot_offset=shot_offset + lo, call=call, skip_shot0=skip_shot0, packed_out=packed_out)
            out[lo:hi] = res[0]
            return res

        futures = []
        for part, nz, (lo, hi) in zip(self.parts, noise.parts_for(self.devices), self._ranges(B)):
            if hi > lo:
                futures.append(((lo, hi), self._pool.submit(run, part, nz, lo, hi)))
        return out, self._gather(futures, shot_offset)

    def sample_device(self, *a, **kw):
        raise NotImplementedError("device-pointer launches address one GPU: use the DeviceProgram of that device (parts[i])")

    def postselect_session(self, *a, **kw):
        return self.parts[0].postselect_session(*a, **kw)  # survivor buffering is sequential by definition: one device

    def level_params(self, component: int, level: int) -> int:
        return self.parts[0].level_params(component, level)

    def evaluate(self, component: int, level: int, params: np.ndarray) -> np.ndarray:
        x = np.asarray(params)
        if x.ndim != 2:
            raise ValueError("params must be 2-D")
        out = np.zeros(x.shape[0], dtype=np.complex64)
        futs = [((lo, hi), self._pool.submit(part.evaluate, component, level, x[lo:hi])) for part, (lo, hi) in zip(self.parts, self._ranges(x.shape[0])) if hi > lo]
        for (lo, hi), fut in futs:
            out[lo:hi] = fut.result()
        return out

    def close(self) -> None:
        for p in self.parts:
            p.close()
