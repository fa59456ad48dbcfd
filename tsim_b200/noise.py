"""Host-side error-mechanism sampler feeding the device sampler.

Mirror of the *sampling* half of the reference's ``ChannelSampler``
(reference ``src/tsim/noise/channels.py:503-658``): per channel, geometric
skips pick the shots in which the channel fires, a conditional CDF picks the
non-identity outcome, and that outcome's precomputed f-pattern is XOR-ed into
the shot's row.  The NumPy ``Generator`` call sequence (one ``geometric`` of
``int(N*p + 7*sigma) + 100`` draws and one ``uniform`` of ``len(positions)``
draws per channel, channels in order) is the same, so for equal seeds the
f-vectors are identical to the reference's -- pinned by
``tests/golden/channel_sampler_*.npz`` (generated from the reference class).

The channel *algebra* (``simplify_channels`` etc., ``channels.py:201-500``)
runs once at compile time inside tsim and is out of scope; use
:meth:`ChannelSampler.from_tsim` to take over an existing tsim sampler's
precomputed tables and RNG, or :meth:`ChannelSampler.from_sparse` /
:meth:`ChannelSampler.from_bit_probs` to build one directly.

Beyond the reference: :meth:`sample_packed` writes the same bits straight into
the 64-bit-word rows the device kernel consumes (8x less host memory traffic
and PCIe volume than ``uint8[B, num_f]``).
"""

from __future__ import annotations

from typing import Any, Sequence

import numpy as np


class ChannelSampler:
    """Geometric-skip sampler over precomputed ``(p_fire, cond_cdf, xor_patterns)`` tables."""

    def __init__(
        self,
        sparse_data: Sequence[tuple[float, np.ndarray, np.ndarray]],
        num_f: int,
        seed: int | None = None,
        *,
        rng: np.random.Generator | None = None,
    ):
        self.num_f = int(num_f)
        self._sparse_data = [
            (float(p), np.asarray(cdf, dtype=np.float64), np.ascontiguousarray(pats, dtype=np.uint8).reshape(-1, self.num_f))
            for p, cdf, pats in sparse_data
        ]
        if rng is None:
            rng = np.random.default_rng(seed if seed is not None else np.random.default_rng().integers(0, 2**30))
        self._rng = rng
        self._words = max(1, (self.num_f + 63) // 64)
        self._packed_patterns = [self._pack(p) for _, _, p in self._sparse_data]

    # -- constructors ---------------------------------------------------------------------------

    @classmethod
    def from_sparse(cls, sparse_data, num_f: int, seed: int | None = None) -> "ChannelSampler":
        return cls(sparse_data, num_f, seed)

    @classmethod
    def from_bit_probs(cls, probs: Sequence[float], seed: int | None = None) -> "ChannelSampler":
        """Independent single-bit channels: f_i fires with probability ``probs[i]``.

        Equals the reference sampler built from ``[error_probs(q) for q in probs]`` with an
        identity ``error_transform`` (channels with ``q <= 1e-15`` are dropped, ``channels.py:607``).
        """
        probs = np.asarray(probs, dtype=np.float64)
        n = len(probs)
        data = []
        for i, q in enumerate(probs):
            p_fire = 1.0 - float(1.0 - q)
            if p_fire <= 1e-15:
                continue
            pat = np.zeros((1, n), dtype=np.uint8)
            pat[0, i] = 1
            data.append((p_fire, np.array([1.0]), pat))
        return cls(data, n, seed)

    @classmethod
    def from_tsim(cls, sampler: Any) -> "ChannelSampler":
        """Adopt a tsim ``ChannelSampler``'s tables and RNG (duck typed; shares the Generator)."""
        num_f = int(sampler.signature_matrix.shape[1])
        return cls(sampler._sparse_data, num_f, rng=sampler._rng)

    # -- sampling -------------------------------------------------------------------------------

    def _pack(self, pats: np.ndarray) -> np.ndarray:
        pad = self._words * 64 - self.num_f
        if pad:
            pats = np.concatenate([pats, np.zeros((pats.shape[0], pad), np.uint8)], axis=1)
        return np.packbits(pats, axis=1, bitorder="little").view(np.uint64).reshape(-1, self._words)

    def _fires(self, num_samples: int):
        """Yield ``(channel, positions, outcome_idx)`` drawing exactly like ``channels.py:638-656``."""
        rng = self._rng
        for ci, (p_fire, cond_cdf, _) in enumerate(self._sparse_data):
            expected = num_samples * p_fire
            sigma = np.sqrt(expected * (1.0 - p_fire))
            n_draws = int(expected + 7.0 * sigma) + 100
            positions = np.cumsum(rng.geometric(p_fire, size=n_draws)) - 1
            positions = positions[positions < num_samples]
            if len(positions) == 0:
                continue
            outcome_idx = np.searchsorted(cond_cdf, rng.uniform(size=len(positions)))
            yield ci, positions, outcome_idx

    def sample(self, num_samples: int = 1) -> np.ndarray:
        """-> ``uint8[num_samples, num_f]`` (0/1), the reference's dense format."""
        result = np.zeros((num_samples, self.num_f), dtype=np.uint8)
        for ci, positions, outcome_idx in self._fires(num_samples):
            result[positions] ^= self._sparse_data[ci][2][outcome_idx]
        return result

    def sample_packed(self, num_samples: int = 1, out: np.ndarray | None = None) -> np.ndarray:
        """Same bits as :meth:`sample`, as ``uint64[num_samples, ceil(num_f/64)]`` little-endian rows."""
        if out is None:
            out = np.zeros((num_samples, self._words), dtype=np.uint64)
        else:
            assert out.shape == (num_samples, self._words) and out.dtype == np.uint64
            out[...] = 0
        for ci, positions, outcome_idx in self._fires(num_samples):
            out[positions] ^= self._packed_patterns[ci][outcome_idx]
        return out

    @property
    def words_per_row(self) -> int:
        return self._words

    def clone_to(self, device: int) -> "DeviceChannelSampler":
        """The same tables, seed and call counter on another GPU."""
        if int(device) == self.device:
            return self
        c = DeviceChannelSampler.__new__(DeviceChannelSampler)
        c.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("_h", "_fin")})
        c.device = int(device)
        c._create()
        return c


def pack_f_rows(f_params: np.ndarray) -> np.ndarray:
    """``uint8/bool[B, num_f]`` -> ``uint64[B, ceil(num_f/64)]`` (bit i of the row = f_i)."""
    f = np.ascontiguousarray(f_params).astype(np.uint8, copy=False)
    B, n = f.shape
    words = max(1, (n + 63) // 64)
    pad = words * 64 - n
    if pad:
        f = np.concatenate([f, np.zeros((B, pad), np.uint8)], axis=1)
    return np.packbits(f, axis=1, bitorder="little").view(np.uint64).reshape(B, words)


class DeviceChannelSampler:
    """Error-mechanism sampler that runs on the GPU (``tsb_noise``; kernel K5).

    Same distribution as :class:`ChannelSampler` / the reference's ``ChannelSampler.sample``
    (``channels.py:624-658``) but a different random stream: like the reference it walks from fire to fire with
    geometric gaps (``:638-656``), here per channel and block of 1024 in-batch shots from a counter-based generator
    keyed by ``(seed, call number)`` -- the rows depend only on (seed, call, in-batch shot index, channel), never on how
    a batch is sliced or sharded -- so parity with the reference is statistical (its own tests use 5-10 % tolerances at 1e5 samples, ``test/unit/noise/test_channels.py:987-1048``).
    With a ``DeviceProgram`` the f rows never leave the GPU (:meth:`DeviceProgram.sample_noisy`).
    """

    def __init__(self, sparse_data, num_f: int, seed: int | None = None, *, device: int = 0):
        import ctypes as C

        from . import _lib

        self.num_f = int(num_f)
        self._words = max(1, (self.num_f + 63) // 64)
        self.seed = int(seed if seed is not None else np.random.default_rng().integers(0, 2**30))
        self.calls = 0  # number of batches drawn so far (part of the generator key)
        self.device = int(device)
        n_out, thr, pats = [], [], []
        for p_fire, cond_cdf, xor_patterns in sparse_data:
            cdf = np.asarray(cond_cdf, dtype=np.float64)
            pat = np.ascontiguousarray(xor_patterns, dtype=np.uint8).reshape(-1, self.num_f)
            if len(cdf) != len(pat) or len(cdf) < 1:
                raise ValueError("cond_cdf and xor_patterns must have one row per non-identity outcome")
            n_out.append(len(cdf))
            for c in cdf:
                thr.append(min(int(float(p_fire) * float(c) * 2.0**64), 2**64 - 1))
            pats.append(pack_f_rows(pat))
        self.n_channels = len(n_out)
        self._n_out = np.asarray(n_out, dtype=np.int32)
        self._thr = np.asarray(thr, dtype=np.uint64)
        self._pat = np.concatenate(pats, axis=0) if pats else np.zeros((0, self._words), np.uint64)
        self._lib = _lib.load()
        self._create()

    def _create(self) -> None:
        import ctypes as C
        import weakref

        from . import _lib

        h = C.c_void_p()
        _lib.check(
            self._lib.tsb_noise_create(
                self.n_channels,
                self._n_out.ctypes.data_as(C.c_void_p),
                self._thr.ctypes.data_as(C.c_void_p),
                np.ascontiguousarray(self._pat).ctypes.data_as(C.c_void_p),
                self._words,
                self.device,
                C.byref(h),
            )
        )
        self._h = h
        self._fin = weakref.finalize(self, self._lib.tsb_noise_destroy, h)

    @classmethod
    def from_bit_probs(cls, probs, seed: int | None = None, *, device: int = 0) -> "DeviceChannelSampler":
        host = ChannelSampler.from_bit_probs(probs, seed=0)
        return cls(host._sparse_data, host.num_f, seed, device=device)

    @classmethod
    def from_host(cls, sampler, seed: int | None = None, *, device: int = 0) -> "DeviceChannelSampler":
        """From a :class:`ChannelSampler` or a tsim ``ChannelSampler`` (uses its precomputed tables)."""
        num_f = int(getattr(sampler, "num_f", None) or sampler.signature_matrix.shape[1])
        return cls(sampler._sparse_data, num_f, seed, device=device)

    @property
    def words_per_row(self) -> int:
        return self._words

    def clone_to(self, device: int) -> "DeviceChannelSampler":
        """The same tables, seed and call counter on another GPU."""
        if int(device) == self.device:
            return self
        c = DeviceChannelSampler.__new__(DeviceChannelSampler)
        c.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("_h", "_fin")})
        c.device = int(device)
        c._create()
        return c

    def next_call(self) -> int:
        c = self.calls
        self.calls += 1
        return c

    def sample_packed(self, num_samples: int = 1, *, shot_offset: int = 0, call: int | None = None, skip_shot0: bool = False) -> np.ndarray:
        """``uint64[num_samples, ceil(num_f/64)]`` rows, generated on the GPU and copied back."""
        import ctypes as C

        from . import _lib

        if call is None:
            call = self.next_call()
        out = np.zeros((num_samples, self._words), dtype=np.uint64)
        _lib.check(
            self._lib.tsb_noise_sample_host(
                self._h, int(num_samples), int(shot_offset), self.seed, int(call), int(skip_shot0), out.ctypes.data_as(C.c_void_p)
            )
        )
        return out

    def sample_device(self, d_f: int, num_samples: int, *, shot_offset: int = 0, call: int | None = None,
                      skip_shot0: bool = False, stream: int = 0) -> None:
        """Write packed rows for in-batch shots ``[shot_offset, shot_offset + num_samples)`` into the device buffer
        at address ``d_f`` (e.g. ``torch.Tensor.data_ptr()``); asynchronous on ``stream``."""
        import ctypes as C

        from . import _lib

        if call is None:
            call = self.next_call()
        _lib.check(
            self._lib.tsb_noise_sample_device(
                self._h, int(num_samples), int(shot_offset), self.seed, int(call), int(skip_shot0), C.c_void_p(d_f),
                C.c_void_p(stream) if stream else None,
            )
        )

    def sample(self, num_samples: int = 1, **kw) -> np.ndarray:
        """Dense ``uint8[num_samples, num_f]`` (the reference's format)."""
        packed = self.sample_packed(num_samples, **kw)
        return np.unpackbits(packed.view(np.uint8), axis=1, bitorder="little", count=self.num_f) if self.num_f else np.zeros((num_samples, 0), np.uint8)


class MultiDeviceChannelSampler:
    """One :class:`DeviceChannelSampler` per GPU with a common seed and call counter.  K5's rows are a pure function of
    (seed, call, in-batch shot index, channel), so the shards of a batch drawn on different devices are the rows a
    single device would have drawn."""

    def __init__(self, sparse_data, num_f: int, seed: int | None = None, *, devices):
        self.seed = int(seed if seed is not None else np.random.default_rng().integers(0, 2**30))
        self.devices = [int(d) for d in devices]
        self.parts = [DeviceChannelSampler(sparse_data, num_f, self.seed, device=d) for d in self.devices]
        self.num_f = int(num_f)
        self.calls = 0

    @classmethod
    def from_bit_probs(cls, probs, seed: int | None = None, *, devices) -> "MultiDeviceChannelSampler":
        host = ChannelSampler.from_bit_probs(probs, seed=0)
        return cls(host._sparse_data, host.num_f, seed, devices=devices)

    @classmethod
    def from_host(cls, sampler, seed: int | None = None, *, devices) -> "MultiDeviceChannelSampler":
        num_f = int(getattr(sampler, "num_f", None) or sampler.signature_matrix.shape[1])
        return cls(sampler._sparse_data, num_f, seed, devices=devices)

    def next_call(self) -> int:
        c = self.calls
        self.calls += 1
        return c

    def parts_for(self, devices):
        if [int(d) for d in devices] != self.devices:
            raise ValueError(f"noise sampler lives on devices {self.devices}, the program on {list(devices)}")
        return self.parts

    def sample_packed(self, num_samples: int = 1, **kw) -> np.ndarray:
        kw.setdefault("call", self.next_call())
        return self.parts[0].sample_packed(num_samples, **kw)

    def sample(self, num_samples: int = 1, **kw) -> np.ndarray:
        kw.setdefault("call", self.next_call())
        return self.parts[0].sample(num_samples, **kw)
