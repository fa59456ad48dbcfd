"""Host-side mirror of tsim's compiled samplers, running the hot path on the B200 library.

Same names, arguments and error behaviour as the reference's ``src/tsim/sampler.py``:

* :func:`sample_program` -- drop-in for ``tsim.sampler.sample_program`` (:117-167); :func:`install`
  rebinds that module-level name when tsim is importable (the seam the reference's own tests patch,
  ``test/unit/test_postselection.py:192-245``).
* :class:`CompiledMeasurementSampler`, :class:`CompiledDetectorSampler`, :class:`CompiledStateProbs`
  -- batching (:340-420), reference sample (:263-276), post-selection (:422-545), output layout
  flags and bit packing (:732-868), ``probability_of`` (:906-953).  They are built from an already
  compiled program plus a channel sampler (``from_tsim`` adopts both from a tsim sampler object),
  because tsim's compile stages stay in tsim.

Nothing here computes amplitudes or draws bits on the CPU: that is ``DeviceProgram`` (CUDA) only.
"""

from __future__ import annotations

import warnings
import weakref
from math import ceil
from typing import Any

import numpy as np

from .backend import DeviceProgram, MultiDeviceProgram, key_words, split_key
from .noise import ChannelSampler, DeviceChannelSampler, MultiDeviceChannelSampler
from .program import CompiledProgram, from_tsim, program_stats
from .shard import pack_bool_rows

_VANISHING = (
    "A vanishing marginal probability distribution was encountered (normalization 0). "
    "This is likely the result of an underflow error. Please report this "
    "as a bug at https://github.com/QuEraComputing/tsim/issues/new."
)

_device_cache: dict[tuple, tuple[Any, Any]] = {}
_DEVICE_CACHE_MAX = 16


def _cache_evict(key) -> None:
    _device_cache.pop(key, None)


def default_devices() -> list[int] | None:
    """``TSIM_B200_DEVICES``: ``"all"`` or a comma-separated list of GPU indices for every handle this module creates
    (``sample_program`` / ``install()`` included), so that a tsim user reaches all GPUs of the box without code changes."""
    import os

    env = os.environ.get("TSIM_B200_DEVICES", "").strip().lower()
    if not env:
        return None
    if env == "all":
        from . import _lib

        return list(range(_lib.load().tsb_device_count()))
    return [int(v) for v in env.split(",") if v.strip()]


def device_program_for(program: Any, *, device: int = 0, devices: Any = None, mode: str = "auto", num_f: int | None = None):
    """Upload ``program`` once per (object, device(s), mode) and reuse the handle on later calls.

    tsim's programs are equinox modules (unhashable), so the cache is keyed on ``id`` and an entry dies with its
    program (``weakref.finalize``) or, for objects that cannot be weakly referenced, when the cache exceeds
    ``_DEVICE_CACHE_MAX`` entries (oldest first)."""
    if isinstance(program, (DeviceProgram, MultiDeviceProgram)):
        return program
    if devices is None:
        devices = default_devices()
    devs = tuple(int(d) for d in devices) if devices is not None else (int(device),)
    key = (id(program), devs, str(mode))
    hit = _device_cache.get(key)
    if hit is not None and (hit[0] is None or hit[0]() is program):
        return hit[1]
    if len(devs) > 1:
        dp = MultiDeviceProgram(from_tsim(program, num_f=num_f), devices=devs, mode=mode)
    else:
        dp = DeviceProgram(from_tsim(program, num_f=num_f), device=devs[0], mode=mode)
    try:
        ref = weakref.ref(program)
        weakref.finalize(program, _cache_evict, key)
    except TypeError:
        ref = None
        dp._pin = program  # keeps id(program) from being reused while the entry lives
    _device_cache[key] = (ref, dp)
    while len(_device_cache) > _DEVICE_CACHE_MAX:
        _device_cache.pop(next(iter(_device_cache)))
    return dp


def check_norm_deviations(devs) -> None:
    """The two thresholds of ``sample_program`` (sampler.py:149-161), per component in order."""
    for dev in devs:
        if np.isclose(dev, 1):
            raise ValueError(_VANISHING)
        if dev > 1e-5:
            warnings.warn(
                "A marginal probability was not normalized correctly "
                f"(normalization deviated from 1 by {dev:.1e}). "
                "This is likely a floating point precision issue.",
                stacklevel=3,
            )


class HostBits(np.ndarray):
    """Result rows in (pinned) host memory.  tsim's ``copy_d2h`` asks an array where it lives before it issues a raw
    ``cudaMemcpy`` on ``unsafe_buffer_pointer()`` (``utils/cuda_helpers.py:32-45, 120-130``): this one answers "cpu"."""

    class _Cpu:
        platform = "cpu"

    def devices(self):
        return {self._Cpu()}


def sample_program(program: Any, f_params: Any, key: Any) -> np.ndarray:
    """Sample all outputs of a compiled program (reference ``sample_program``, sampler.py:117-167).

    ``program``: tsim / tsim_b200 ``CompiledProgram`` or a ``DeviceProgram``; ``f_params``: array
    ``[batch, num_f]`` of 0/1; ``key``: jax PRNG key or ``(k0, k1)``.  Returns ``bool[batch, num_outputs]``.
    """
    f = np.asarray(f_params)
    # the f vector may be wider than the highest index the program references (reference: simply ignored)
    dp = device_program_for(program, num_f=f.shape[1] if f.ndim == 2 else None)
    if dp.num_outputs == 0:
        return np.zeros((f.shape[0], 0), dtype=np.bool_)
    bits, devs = dp.sample(f, key)
    check_norm_deviations(devs)
    return bits.view(HostBits)


class _JnpShim:
    """``jax.numpy`` as seen by ``tsim.sampler`` after :func:`install`: host results stay on the host when
    ``_sample_batches`` concatenates its batches (sampler.py:411) instead of being uploaded by ``jnp.concatenate``."""

    def __init__(self, jnp):
        self._jnp = jnp

    def __getattr__(self, name):
        return getattr(self._jnp, name)

    def concatenate(self, arrays, axis=0, **kw):
        arrays = list(arrays)
        if arrays and all(isinstance(a, np.ndarray) for a in arrays):
            return np.concatenate(arrays, axis=axis).view(HostBits)
        return self._jnp.concatenate(arrays, axis=axis, **kw)


_installed: dict[str, Any] = {}


def install() -> bool:
    """Rebind ``tsim.sampler.sample_program`` to the B200 backend.  False if tsim is not importable.

    The reference's ``_sample_batches`` then hands our host-resident result to ``jnp.concatenate`` and ``copy_d2h``
    (sampler.py:411-413); both names are rebound inside ``tsim.sampler`` as well so that host rows pass through
    untouched (no re-upload, no raw ``cudaMemcpy`` on a host pointer).  :func:`uninstall` restores all three."""
    try:
        import tsim.sampler as ts
    except Exception:
        return False
    if _installed:
        return True
    _installed.update(sample_program=ts.sample_program, copy_d2h=getattr(ts, "copy_d2h", None), jnp=getattr(ts, "jnp", None))
    ts.sample_program = sample_program
    ref_copy = _installed["copy_d2h"]
    if ref_copy is not None:

        def copy_d2h(src, *, dst=None):
            if isinstance(src, np.ndarray):  # already on the host (page-locked pool of the backend)
                if dst is None:
                    return np.asarray(src)
                dst[...] = src
                return dst
            return ref_copy(src, dst=dst)

        ts.copy_d2h = copy_d2h
    if _installed["jnp"] is not None:
        ts.jnp = _JnpShim(_installed["jnp"])
    return True


def uninstall() -> None:
    """Undo :func:`install`."""
    if not _installed:
        return
    import tsim.sampler as ts

    ts.sample_program = _installed["sample_program"]
    if _installed["copy_d2h"] is not None:
        ts.copy_d2h = _installed["copy_d2h"]
    if _installed["jnp"] is not None:
        ts.jnp = _installed["jnp"]
    _installed.clear()


class _CompiledSamplerBase:
    """Reference ``_CompiledSamplerBase`` (sampler.py:170-609) on top of a ``DeviceProgram``."""

    #: upper bound for automatically chosen batches (the reference derives one from free memory,
    #: sampler.py:308-320; here a shot needs 8*(ceil(num_f/64)+ceil(n_out/64)) bytes of HBM)
    MAX_AUTO_BATCH = 1 << 22

    def __init__(
        self,
        program: Any,
        channel_sampler: ChannelSampler,
        *,
        num_detectors: int | None = None,
        seed: int | None = None,
        key: tuple[int, int] | None = None,
        device: int = 0,
        devices: Any = None,
        mode: str = "auto",
        joint: bool = False,
    ):
        """``devices``: GPUs of this box to shard every batch over, from this one process (default: ``device`` alone).  The
        sampled bits do not depend on it (RNG counters are in-batch shot indices)."""
        if seed is None and key is None:
            seed = int(np.random.default_rng().integers(0, 2**30))
        # jax.random.key(seed) (sampler.py:198)
        self._key = key_words(key) if key is not None else ((int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF)
        self._program: CompiledProgram = from_tsim(program, num_f=getattr(channel_sampler, "num_f", None))
        if devices is None:
            devices = default_devices()
        if devices is not None and len(list(devices)) > 1:
            self._device_program = MultiDeviceProgram(self._program, devices=list(devices), mode=mode, joint=joint)
            if isinstance(channel_sampler, DeviceChannelSampler):  # the same tables and seed on every device
                ds = channel_sampler
                channel_sampler = MultiDeviceChannelSampler.__new__(MultiDeviceChannelSampler)
                channel_sampler.seed, channel_sampler.devices, channel_sampler.num_f, channel_sampler.calls = ds.seed, [int(d) for d in devices], ds.num_f, ds.calls
                channel_sampler.parts = [ds.clone_to(d) for d in channel_sampler.devices]
        else:
            if devices is not None and len(list(devices)) == 1:
                device = int(list(devices)[0])
            self._device_program = DeviceProgram(self._program, device=device, mode=mode, joint=joint)
        self._channel_sampler = channel_sampler
        self._num_detectors = int(self._program.num_detectors if num_detectors is None else num_detectors)

        prog = self._program
        self._direct_f_indices = np.asarray(prog.direct_f_indices)
        self._direct_flips = np.asarray(prog.direct_flips, dtype=np.bool_)
        self._direct_reindex = np.asarray(prog.output_reindex) if prog.output_reindex is not None else None
        n_direct = len(self._direct_f_indices)
        self._direct_zero_copy = (
            n_direct > 0
            and self._direct_reindex is None
            and not self._direct_flips.any()
            and np.array_equal(self._direct_f_indices, np.arange(n_direct))
        )
        self._direct_global_indices = np.asarray(prog.output_order[:n_direct], dtype=np.int32)
        self._direct_output_mask = np.zeros(prog.num_outputs, dtype=np.bool_)
        if n_direct > 0:
            self._direct_output_mask[self._direct_global_indices] = True
        self._direct_detector_mask = self._direct_output_mask[: self._num_detectors].copy()

    @classmethod
    def from_tsim(cls, sampler: Any, **kw):
        """Adopt a tsim sampler's compiled program, channel sampler (shared RNG) and key."""
        import jax

        key = tuple(int(v) for v in np.asarray(jax.random.key_data(sampler._key)).reshape(2))
        return cls(
            sampler._program,
            ChannelSampler.from_tsim(sampler._channel_sampler),
            num_detectors=sampler._num_detectors,
            key=key,
            **kw,
        )

    # -- helpers mirrored from the reference ---------------------------------------------------
    def _next_subkey(self) -> tuple[int, int]:
        self._key, sub = split_key(self._key)  # sampler.py:399
        return sub

    def _run(self, f_params_np: np.ndarray, *, packed: bool = False) -> np.ndarray:
        if packed:
            bits, devs = self._device_program.sample(f_params_np, self._next_subkey(), packed_out=True)
            check_norm_deviations(devs)
            return bits
        # module-level name looked up at call time, like the reference (sampler.py:274,400,484)
        return sample_program(self._device_program, f_params_np, self._next_subkey())

    def _compute_direct_outputs(self, f_params_np: np.ndarray) -> np.ndarray:
        batch = f_params_np.shape[0]
        num_outputs = self._program.num_outputs
        n_direct = len(self._direct_f_indices)
        if n_direct == 0:
            return np.zeros((batch, num_outputs), dtype=np.bool_)
        raw = (f_params_np[:, self._direct_f_indices].astype(np.bool_)) ^ self._direct_flips
        out = np.zeros((batch, num_outputs), dtype=np.bool_)
        out[:, self._direct_global_indices] = raw
        return out

    def _compute_reference_sample(self) -> np.ndarray:
        num_f = self._channel_sampler.num_f
        f_ref = np.zeros((1, num_f), dtype=np.uint8)
        if not self._program.components:
            return self._compute_direct_outputs(f_ref)[0]
        return np.asarray(self._run(f_ref)[0], dtype=np.bool_)

    def _peak_bytes_per_sample(self) -> int:
        """Device bytes per shot of one batch (reference ``_peak_bytes_per_sample``, sampler.py:294-306, estimates XLA's
        intermediates; here: packed f and output rows, their byte-format staging and the bit-sliced scratch)."""
        prog = self._program
        num_f, n_out = prog.infer_num_f(), int(prog.num_outputs)
        total_f = sum(len(c.f_selection) for c in prog.components)
        n_draws = sum(len(c.compiled_scalar_graphs) - 1 for c in prog.components)
        return max(1, 8 * (-(-num_f // 64) + -(-n_out // 64)) + num_f + n_out + (total_f + n_draws + 32 + 7) // 8 + 4)

    def _estimate_batch_size(self) -> int:
        """Largest batch that fits in half of the free device memory (reference sampler.py:308-320), capped at
        ``MAX_AUTO_BATCH``: the device pipeline works through a batch in slices, so larger batches buy nothing."""
        cached = getattr(self, "_auto_batch", None)
        if cached is not None:
            return cached
        mem_info = getattr(self._device_program, "mem_info", None)
        if mem_info is None:
            return self.MAX_AUTO_BATCH
        # asked once per sampler: cudaMemGetInfo costs up to milliseconds, a sample() call of 10^6 shots less than one
        free, _total = mem_info()
        self._auto_batch = max(1, min(self.MAX_AUTO_BATCH, int(free * 0.5) // self._peak_bytes_per_sample()))
        return self._auto_batch

    def _resolve_batch_size(self, shots: int, batch_size: int | None, *, compute_reference: bool) -> int:
        if batch_size is None:
            max_batch_size = self._estimate_batch_size()
            num_batches = max(1, ceil(shots / max_batch_size))
            batch_size = ceil(shots / num_batches)
        if compute_reference and batch_size * ceil(shots / batch_size) == shots:
            batch_size += 1
        return batch_size

    def _sample_direct(self, shots: int) -> np.ndarray:
        f_params = self._channel_sampler.sample(shots)
        result = f_params[:, self._direct_f_indices] ^ self._direct_flips
        if self._direct_reindex is not None:
            result = result[:, self._direct_reindex]
        return result.view(np.bool_)

    def _sample_batches(self, shots: int, batch_size: int | None = None, *, compute_reference: bool = False,
                        packed: bool = False):
        """Reference ``_sample_batches`` (sampler.py:340-420).  ``packed=True`` (not in the reference) keeps the
        result as the device's ``uint64[shots, ceil(n_out/64)]`` rows; the reference row is still ``bool[n_out]``."""
        # direct-only programs follow the reference's host shortcut (sampler.py:365-370) unless the noise itself lives on
        # the device: then K5 -> direct gather -> packed rows never materialise the f matrix on the host
        on_device = isinstance(self._channel_sampler, (DeviceChannelSampler, MultiDeviceChannelSampler))
        host_direct = not self._program.components and not on_device
        if packed and (shots == 0 or host_direct):
            res = self._sample_batches(shots, batch_size, compute_reference=compute_reference)
            if compute_reference:
                return pack_bool_rows(res[0]), res[1]
            return pack_bool_rows(res)
        if shots < 0:
            raise ValueError(f"shots must be non-negative, got {shots}")
        if batch_size is not None and batch_size < 1:
            raise ValueError(f"batch_size must be at least 1, got {batch_size}")
        if shots == 0:
            empty = np.empty((0, self._program.num_outputs), dtype=np.bool_)
            if compute_reference:
                return empty, np.zeros(self._program.num_outputs, dtype=np.bool_)
            return empty
        if host_direct:
            samples = self._sample_direct(shots)
            if compute_reference:
                return samples, self._compute_reference_sample()
            return samples
        if batch_size is None:
            max_batch_size = self._estimate_batch_size()
            num_batches = max(1, ceil(shots / max_batch_size))
            batch_size = ceil(shots / num_batches)
        else:
            num_batches = ceil(shots / batch_size)
        if compute_reference and batch_size * num_batches == shots:
            batch_size += 1

        batches = []
        reference = None
        packed_host = hasattr(self._channel_sampler, "sample_packed")
        for _ in range(num_batches):
            if on_device:
                # noise, sampling and packing all on the GPU; the f rows never reach the host
                samples, devs = self._device_program.sample_noisy(
                    self._channel_sampler, batch_size, self._next_subkey(), skip_shot0=compute_reference and reference is None,
                    packed_out=packed,
                )
                check_norm_deviations(devs)
            else:
                # same bits either way; packed rows move 8x fewer bytes than the reference's uint8 matrix
                f_params_np = (
                    self._channel_sampler.sample_packed(batch_size) if packed_host else self._channel_sampler.sample(batch_size)
                )
                if compute_reference and reference is None:
                    f_params_np[0] = 0
                samples = self._run(f_params_np, packed=packed)
            if compute_reference and reference is None:
                reference = np.asarray(samples[0]).copy()
                if packed:
                    reference = np.unpackbits(reference.view(np.uint8), bitorder="little", count=self._program.num_outputs).astype(np.bool_)
                samples = samples[1:]
            batches.append(samples)
        result = (batches[0] if len(batches) == 1 else np.concatenate(batches, axis=0))[:shots]
        if compute_reference:
            return result, reference
        return result

    def _sample_batches_with_postselection(
        self, shots: int, batch_size: int | None, *, postselection_mask: np.ndarray, compute_reference: bool = False,
        xor_detector_ref: bool = False,
    ):
        """Reference ``_sample_batches_with_postselection`` (sampler.py:422-545).  With a real ``DeviceProgram`` the
        survivor buffering runs on the GPU (``tsb_postselect``: same chunks, same batches, same key schedule, hence the
        same bits); ``TSIM_B200_POSTSELECT=host`` or a device object without sessions keeps the host buffering."""
        import os

        if (
            shots > 0
            and self._program.components
            and hasattr(self._device_program, "postselect_session")
            and os.environ.get("TSIM_B200_POSTSELECT", "device") != "host"
        ):
            return self._postselect_device(shots, batch_size, postselection_mask=postselection_mask,
                                           compute_reference=compute_reference, xor_detector_ref=xor_detector_ref)
        return self._postselect_host(shots, batch_size, postselection_mask=postselection_mask,
                                     compute_reference=compute_reference, xor_detector_ref=xor_detector_ref)

    def _postselect_device(self, shots: int, batch_size: int | None, *, postselection_mask: np.ndarray,
                           compute_reference: bool = False, xor_detector_ref: bool = False):
        if batch_size is not None and batch_size < 1:
            raise ValueError(f"batch_size must be at least 1, got {batch_size}")
        nd = self._num_detectors
        num_outputs = self._program.num_outputs
        postselect_direct = postselection_mask & self._direct_detector_mask
        if batch_size is None:
            batch_size = self._resolve_batch_size(shots, batch_size, compute_reference=False)
        reference = self._compute_reference_sample() if compute_reference else None
        use_ref = xor_detector_ref and reference is not None
        session = self._device_program.postselect_session(
            shots, batch_size, postselect_direct, reference[:nd] if use_ref else None, nd
        )
        on_device = isinstance(self._channel_sampler, DeviceChannelSampler)  # multi-device noise: rows come back through the host
        packed_host = hasattr(self._channel_sampler, "sample_packed")
        shot_idx = 0
        pending = 0
        while shot_idx < shots:
            chunk = min(batch_size, shots - shot_idx)
            if on_device:
                pending = session.push_noise(self._channel_sampler, chunk)
            else:
                f = self._channel_sampler.sample_packed(chunk) if packed_host else self._channel_sampler.sample(chunk)
                pending = session.push_host(f)
            shot_idx += chunk
            while pending >= batch_size:  # _flush (sampler.py:494-498): one key per dispatched batch
                pending, devs = session.dispatch(self._next_subkey())
                check_norm_deviations(devs)
        if pending:  # _flush(final=True): padded partial batch
            pending, devs = session.dispatch(self._next_subkey(), final=True)
            check_norm_deviations(devs)
        xk = xd = None
        if use_ref:
            xk = np.zeros(num_outputs, dtype=np.bool_)
            xk[:nd] = reference[:nd]
            xd = np.zeros(num_outputs, dtype=np.bool_)
            xd[:nd] = reference[:nd] & self._direct_detector_mask
        result, was_discarded = session.finish(xk, xd)
        if compute_reference:
            return result, reference, was_discarded
        return result, None, was_discarded

    def _postselect_host(
        self, shots: int, batch_size: int | None, *, postselection_mask: np.ndarray, compute_reference: bool = False,
        xor_detector_ref: bool = False,
    ):
        if shots < 0:
            raise ValueError(f"shots must be non-negative, got {shots}")
        if batch_size is not None and batch_size < 1:
            raise ValueError(f"batch_size must be at least 1, got {batch_size}")
        num_outputs = self._program.num_outputs
        nd = self._num_detectors
        if shots == 0:
            empty = np.empty((0, num_outputs), dtype=np.bool_)
            none_discarded = np.empty(0, dtype=np.bool_)
            if compute_reference:
                return empty, np.zeros(num_outputs, dtype=np.bool_), none_discarded
            return empty, None, none_discarded
        postselect_direct = postselection_mask & self._direct_detector_mask
        if not self._program.components:
            samples = self._sample_direct(shots)
            if compute_reference:
                reference = self._compute_reference_sample()
                if xor_detector_ref:
                    samples[:, :nd] ^= reference[:nd]
                return samples, reference, np.zeros(shots, dtype=np.bool_)
            return samples, None, np.zeros(shots, dtype=np.bool_)
        if batch_size is None:
            batch_size = self._resolve_batch_size(shots, batch_size, compute_reference=False)
        reference = self._compute_reference_sample() if compute_reference else None

        result = np.zeros((shots, num_outputs), dtype=np.bool_)
        was_discarded = np.zeros(shots, dtype=np.bool_)
        pending_f: list[np.ndarray] = []  # survivor rows waiting for a full device batch
        pending_idx: list[np.ndarray] = []
        n_pending = 0

        def dispatch(f_batch: np.ndarray, indices: np.ndarray, n_valid: int) -> None:
            out = self._run(f_batch)
            result[indices[:n_valid]] = out[:n_valid]

        def flush(final: bool = False) -> None:
            nonlocal pending_f, pending_idx, n_pending
            if n_pending >= batch_size or (final and n_pending):
                f_all = np.concatenate(pending_f, axis=0)
                i_all = np.concatenate(pending_idx)
                pos = 0
                while n_pending - pos >= batch_size:
                    dispatch(f_all[pos : pos + batch_size], i_all[pos : pos + batch_size], batch_size)
                    pos += batch_size
                if final and pos < n_pending:
                    n_valid = n_pending - pos
                    f_batch = np.empty((batch_size, f_all.shape[1]), dtype=f_all.dtype)
                    f_batch[:n_valid] = f_all[pos:]
                    f_batch[n_valid:] = f_all[pos]  # fixed batch shape, padding rows discarded (sampler.py:499-505)
                    dispatch(f_batch, i_all[pos:], n_valid)
                    pos = n_pending
                pending_f = [f_all[pos:]] if pos < n_pending else []
                pending_idx = [i_all[pos:]] if pos < n_pending else []
                n_pending -= pos

        shot_idx = 0
        while shot_idx < shots:
            chunk = min(batch_size, shots - shot_idx)
            f_params_np = self._channel_sampler.sample(chunk)
            direct_full = self._compute_direct_outputs(f_params_np)
            det_cols = direct_full[:, :nd]
            if xor_detector_ref and reference is not None:
                det_cols = det_cols ^ reference[:nd]
            discarded = (det_cols & postselect_direct).any(axis=1)
            result[shot_idx : shot_idx + chunk, :nd] = direct_full[:, :nd]
            was_discarded[shot_idx : shot_idx + chunk] = discarded
            survivors = np.flatnonzero(~discarded)
            if survivors.size:
                pending_f.append(f_params_np[survivors])
                pending_idx.append(shot_idx + survivors)
                n_pending += survivors.size
            shot_idx += chunk
            flush()
        flush(final=True)

        if xor_detector_ref and reference is not None:
            det_ref = reference[:nd]
            result[~was_discarded, :nd] ^= det_ref
            result[was_discarded, :nd] ^= det_ref & self._direct_detector_mask
        if compute_reference:
            return result, reference, was_discarded
        return result, None, was_discarded

    def __repr__(self) -> str:
        s = program_stats(self._program)
        info = self._device_program.info
        return (
            f"{type(self).__name__}({s['direct']} direct, {s['graphs']} graphs, "
            f"{s['max_outputs_per_component']} outputs for largest cc, ≤ {s['max_params']} parameters, "
            f"{s['A_terms']} A terms, {s['B_terms']} B terms, {s['C_terms']} C terms, {s['D_terms']} D terms, "
            f"{info['data_bytes']} B packed on cuda:{self._device_program.device}, "
            f"{'resident' if info['resident'] else 'streamed'}, mode={('faithful', 'fast', 'sliced')[info['mode']]})"
        )


class CompiledMeasurementSampler(_CompiledSamplerBase):
    """Reference ``CompiledMeasurementSampler`` (sampler.py:612-662)."""

    def sample(self, shots: int, *, batch_size: int | None = None) -> np.ndarray:
        return self._sample_batches(shots, batch_size)


def _layout_segments(nd: int, n_out: int, *, prepend: bool, append: bool, separate: bool):
    """Column ranges of ``CompiledDetectorSampler.sample``'s result (sampler.py:852-868) over the combined columns
    ``[detectors | observables]``."""
    det, obs = (0, nd), (nd, n_out - nd)
    if prepend and append:
        return [obs, det, obs]
    if append or separate:
        return [det, obs]
    if prepend:
        return [obs, det]
    return [det]


def _packed_columns(rows: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Columns ``[lo, hi)`` of packed ``uint64[B, W]`` rows as ``np.packbits(..., bitorder="little")`` bytes.

    Up to 64 columns (the detector / observable split of a typical program) are cut out with contiguous word
    operations and one narrowing cast; wider ranges take the word-by-word form."""
    B, W = rows.shape
    n = hi - lo
    nb = (n + 7) // 8
    if n <= 0:
        return np.zeros((B, 0), dtype=np.uint8)
    if n > 64:
        return _packed_columns_words(rows, lo, hi)
    w, sh = divmod(lo, 64)
    v = rows[:, w] >> np.uint64(sh) if sh else rows[:, w]
    if sh and sh + n > 64 and w + 1 < W:
        v = v | (rows[:, w + 1] << np.uint64(64 - sh))
    if n < 64:
        v = v & np.uint64((1 << n) - 1)
    width = 1 if nb <= 1 else 2 if nb <= 2 else 4 if nb <= 4 else 8
    out = v.astype({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[width]).view(np.uint8).reshape(B, width)
    return out if nb == width else np.ascontiguousarray(out[:, :nb])


def _packed_columns_words(rows: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """General form of ``_packed_columns`` for column ranges that straddle more than eight bytes."""
    B, W = rows.shape
    n = hi - lo
    n_words = max(1, (n + 63) // 64)
    out = np.zeros((B, n_words), dtype=np.uint64)
    for k in range(n_words if n > 0 else 0):
        off = lo + 64 * k
        w, sh = divmod(off, 64)
        if w >= W:
            break
        v = rows[:, w] >> np.uint64(sh)
        if sh and w + 1 < W:
            v = v | (rows[:, w + 1] << np.uint64(64 - sh))
        out[:, k] = v
    tail = n - 64 * (n_words - 1)
    if 0 <= tail < 64:
        out[:, n_words - 1] &= np.uint64((1 << tail) - 1)
    return np.ascontiguousarray(out.view(np.uint8)[:, : (n + 7) // 8])


def _concat_packed(parts: list[tuple[np.ndarray, int]]) -> np.ndarray:
    """Concatenate bit-packed column blocks ``(bytes[B, ceil(n/8)], n)`` along the bit axis."""
    if len(parts) == 1:
        return parts[0][0]
    bits = [np.unpackbits(a, axis=1, bitorder="little", count=n) for a, n in parts]
    return np.packbits(np.concatenate(bits, axis=1), axis=1, bitorder="little")


def _maybe_bit_pack(array: np.ndarray, *, bit_packed: bool) -> np.ndarray:
    if not bit_packed:
        return array
    return np.packbits(array.astype(np.bool_), axis=1, bitorder="little")


class CompiledDetectorSampler(_CompiledSamplerBase):
    """Reference ``CompiledDetectorSampler`` (sampler.py:672-868)."""

    #: with a device channel sampler the layout flags are applied on the GPU (False: host NumPy, same bits)
    DEVICE_LAYOUT = True

    def sample(
        self,
        shots: int,
        *,
        batch_size: int | None = None,
        prepend_observables: bool = False,
        append_observables: bool = False,
        separate_observables: bool = False,
        bit_packed: bool = False,
        use_detector_reference_sample: bool = False,
        use_observable_reference_sample: bool = False,
        postselection_mask: np.ndarray | None = None,
    ):
        if separate_observables and (prepend_observables or append_observables):
            raise ValueError(
                "Can't specify separate_observables=True with append_observables=True or prepend_observables=True"
            )
        compute_reference = use_detector_reference_sample or use_observable_reference_sample
        nd = self._num_detectors
        if postselection_mask is not None:
            mask = np.asarray(postselection_mask, dtype=np.bool_)
            if mask.shape != (nd,):
                raise ValueError(f"postselection_mask must have shape ({nd},), got {mask.shape}")
            postselection_mask = mask
            if not (mask & self._direct_detector_mask).any() or not self._program.components:
                postselection_mask = None

        if (self.DEVICE_LAYOUT and postselection_mask is None and shots > 0 and isinstance(self._channel_sampler, DeviceChannelSampler)
                and isinstance(self._device_program, DeviceProgram) and (batch_size is None or batch_size >= 1)):
            # noise, sampling, column selection, reference XOR and bit packing all on the GPU: only the bytes of the
            # returned arrays cross PCIe (tsb_sample_noisy_host_layout)
            return self._sample_layout_device(
                shots, batch_size, prepend=prepend_observables, append=append_observables, separate=separate_observables,
                bit_packed=bit_packed, ref_det=use_detector_reference_sample, ref_obs=use_observable_reference_sample,
            )
        if postselection_mask is not None:
            if compute_reference:
                samples, reference, direct_discarded = self._sample_batches_with_postselection(
                    shots, batch_size, postselection_mask=postselection_mask, compute_reference=True,
                    xor_detector_ref=use_detector_reference_sample,
                )
                if use_observable_reference_sample:
                    samples[~direct_discarded, nd:] ^= reference[nd:]
            else:
                samples, _, _ = self._sample_batches_with_postselection(shots, batch_size, postselection_mask=postselection_mask)
        elif bit_packed and shots > 0 and (self._program.components or isinstance(self._channel_sampler, (DeviceChannelSampler, MultiDeviceChannelSampler))):
            # packed end to end: the device's uint64 rows are sliced with word shifts, never expanded to bools
            n_out = self._program.num_outputs
            if compute_reference:
                rows, reference = self._sample_batches(shots, batch_size, compute_reference=True, packed=True)
                ref = reference.copy()
                if not use_detector_reference_sample:
                    ref[:nd] = False
                if not use_observable_reference_sample:
                    ref[nd:] = False
                rows = rows ^ pack_bool_rows(ref[None, :])
            else:
                rows = self._sample_batches(shots, batch_size, packed=True)
            det, obs = (lambda: _packed_columns(rows, 0, nd)), (lambda: _packed_columns(rows, nd, n_out))
            if prepend_observables and append_observables:
                return _concat_packed([(obs(), n_out - nd), (det(), nd), (obs(), n_out - nd)])
            if append_observables:
                return _packed_columns(rows, 0, n_out)
            if prepend_observables:
                return _concat_packed([(obs(), n_out - nd), (det(), nd)])
            if separate_observables:
                return det(), obs()
            return det()
        elif compute_reference:
            samples, reference = self._sample_batches(shots, batch_size, compute_reference=True)
            samples = np.array(samples, copy=True)
            if use_detector_reference_sample:
                samples[:, :nd] ^= reference[:nd]
            if use_observable_reference_sample:
                samples[:, nd:] ^= reference[nd:]
        else:
            samples = self._sample_batches(shots, batch_size)

        det_samples = samples[:, :nd]
        obs_samples = samples[:, nd:]
        if prepend_observables and append_observables:
            return _maybe_bit_pack(np.concatenate([obs_samples, det_samples, obs_samples], axis=1), bit_packed=bit_packed)
        if append_observables:
            return _maybe_bit_pack(samples, bit_packed=bit_packed)
        if prepend_observables:
            return _maybe_bit_pack(np.concatenate([obs_samples, det_samples], axis=1), bit_packed=bit_packed)
        if separate_observables:
            return _maybe_bit_pack(det_samples, bit_packed=bit_packed), _maybe_bit_pack(obs_samples, bit_packed=bit_packed)
        return _maybe_bit_pack(det_samples, bit_packed=bit_packed)


    def _sample_layout_device(self, shots: int, batch_size: int | None, *, prepend: bool, append: bool, separate: bool,
                              bit_packed: bool, ref_det: bool, ref_obs: bool):
        """The batch loop of ``_sample_batches`` (sampler.py:340-420) with the layout flags applied on the device."""
        nd, n_out = self._num_detectors, self._program.num_outputs
        compute_reference = ref_det or ref_obs
        if batch_size is None:
            max_batch_size = self._estimate_batch_size()
            num_batches = max(1, ceil(shots / max_batch_size))
            batch_size = ceil(shots / num_batches)
        else:
            num_batches = ceil(shots / batch_size)
        if compute_reference and batch_size * num_batches == shots:
            batch_size += 1
        segments = _layout_segments(nd, n_out, prepend=prepend, append=append, separate=separate)
        split = 1 if separate else 0  # separate_observables: detectors and observables as two arrays
        dp = self._device_program
        ref_mask = None
        if compute_reference:
            m = np.zeros(n_out, dtype=np.bool_)
            m[:nd] = ref_det
            m[nd:] = ref_obs
            ref_mask = pack_bool_rows(m[None, :])[0]
        from .backend import _result_pool

        total = num_batches * batch_size - (1 if compute_reference else 0)

        def take(row_bytes):
            return _result_pool.take((total, row_bytes), np.uint8) if total * row_bytes > 0 else np.empty((total, row_bytes), np.uint8)

        rb = dp.layout_row_bytes(segments, bit_packed=bit_packed, split=split)
        out, out2 = take(rb[0]), (take(rb[1]) if separate else None)
        pos, xor_row = 0, None
        for b in range(num_batches):
            first_ref = compute_reference and b == 0
            part, _, devs, row0 = dp.sample_noisy_layout(
                self._channel_sampler, batch_size, self._next_subkey(), segments, bit_packed=bit_packed, split=split,
                xor_row=xor_row, skip_shot0=first_ref, ref_mask=ref_mask if first_ref else None, out=out[pos:],
                out2=out2[pos:] if out2 is not None else None,
            )
            pos += part.shape[0]
            check_norm_deviations(devs)
            if first_ref:
                xor_row = row0 & ref_mask

        def finish(a):
            a = a[:shots]
            return a if bit_packed else a.view(np.bool_)

        if separate:
            return finish(out), finish(out2)
        return finish(out)


class CompiledStateProbs(_CompiledSamplerBase):
    """Reference ``CompiledStateProbs`` (sampler.py:871-953): joint-mode programs, two levels per component."""

    def __init__(self, program, channel_sampler, **kw):
        super().__init__(program, channel_sampler, joint=True, **kw)

    def probability_of(self, state: np.ndarray, *, batch_size: int) -> np.ndarray:
        if batch_size < 1:
            raise ValueError(f"batch_size must be at least 1, got {batch_size}")
        prog = self._program
        state = np.asarray(state)
        if state.shape != (prog.num_outputs,):
            raise ValueError(f"state must have shape ({prog.num_outputs},), got {state.shape}")
        f_samples = self._channel_sampler.sample(batch_size)
        p_norm = np.ones(batch_size, dtype=np.float32)
        p_joint = np.ones(batch_size, dtype=np.float32)
        n_direct = len(prog.direct_f_indices)
        if n_direct > 0:
            direct_bits = f_samples[:, prog.direct_f_indices].astype(np.bool_) ^ prog.direct_flips
            targets = state[prog.output_order[:n_direct]].astype(np.bool_)
            p_joint = p_joint * (direct_bits == targets).all(axis=1)
        for ci, component in enumerate(prog.components):
            assert len(component.compiled_scalar_graphs) == 2
            f_selected = f_samples[:, component.f_selection]
            p_norm = p_norm * np.abs(self._device_program.evaluate(ci, 0, f_selected))
            component_state = state[list(component.output_indices)]
            joint_params = np.hstack([f_selected, np.tile(component_state, (batch_size, 1))])
            p_joint = p_joint * np.abs(self._device_program.evaluate(ci, 1, joint_params))
        with np.errstate(all="ignore"):
            return np.asarray(p_joint / p_norm)
