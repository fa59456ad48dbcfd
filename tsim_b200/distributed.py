"""Shot-sharded sampling over the GPUs of one node.

One process per GPU (``torchrun``); every rank uploads the same program and builds the same device channel
sampler.  A batch of ``B`` shots is cut into contiguous, balanced row ranges (``shard.shard_range``); rank ``r``
generates the f rows of its range on its GPU (K5 -- rows depend only on (seed, call, in-batch shot index), so no
host scatter is needed), samples them with ``shot_offset = lo`` (the RNG counter is the in-batch shot index) and
takes part in ONE ``all_gather`` of the packed output rows.  The gathered bits are identical to a single-GPU run
of the same batch.  ``torch.distributed`` (NCCL) is plumbing; there is no other collective on the path.
"""

from __future__ import annotations

from math import ceil

import numpy as np

from .backend import DeviceProgram, split_key
from .noise import DeviceChannelSampler
from .shard import gather_packed_rows, shard_range


def nccl_defaults() -> None:
    """Environment defaults for the one collective of the path; call before the process group comes up.

    The sampling kernel is one wave of CTAs that fill their SMs (896 threads x 72 registers, 227 KB of shared memory); its
    launch plan leaves 8 SMs free.  Eight NCCL channels make the all-gather fit into those SMs so that it overlaps the
    next step's kernel instead of queueing behind it (N = 8: 7.0e9 -> 7.7e9 shots/s, profiles/).  ``setdefault``: an
    explicit NCCL_MAX_NCHANNELS of the user wins."""
    import os

    os.environ.setdefault("NCCL_MAX_NCHANNELS", "8")


class PeerGather:
    """The gather of packed output rows over NVLink peer memory, without SMs.

    Every rank owns ``n_buffers`` receive buffers ``int64[world * shots, words]`` in symmetric memory
    (``torch.distributed._symmetric_memory``: same allocation on every rank, peer-mapped through CUDA IPC / fabric
    handles).  The sampling pipeline writes a rank's rows straight into its slice of its own buffer
    (:meth:`local_rows`); :meth:`push` then copies that slice into the same slice of every peer's buffer with
    copy-engine transfers (``tsb_memcpy_peer_async``) on a side stream and joins a signal-pad barrier -- when the
    returned event fires, every rank's rows have landed here.  No SM takes part, so the exchange of step i hides
    completely behind the sampling kernel of step i + 1, which fills its SMs with one 228 KB CTA each and leaves an
    NCCL all-gather only the eight SMs its launch plan spares (N = 8: 0.81 -> 0.95 ms per step with NCCL).
    Raises if symmetric memory cannot be set up (callers fall back to ``all_gather_into_tensor``)."""

    def __init__(self, shots: int, words: int, device: int, n_buffers: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.shots, self.words = int(shots), int(words)
        dev = torch.device("cuda", device)
        self._torch, self._lib = torch, _lib.load()
        self._check = _lib.check
        self.bufs = [symm_mem.empty((self.world * self.shots, self.words), dtype=torch.int64, device=dev) for _ in range(n_buffers)]
        self.hdls = [symm_mem.rendezvous(b, dist.group.WORLD) for b in self.bufs]
        self.side = torch.cuda.Stream(dev)
        self._consumed = [None] * n_buffers
        self._slice_bytes = self.shots * self.words * 8
        for h in self.hdls:
            if len(h.buffer_ptrs) != self.world:
                raise RuntimeError("symmetric memory rendezvous did not return one buffer per rank")

    def local_rows(self, i: int):
        """This rank's slice of receive buffer ``i``: the destination of the sampling pipeline's output rows."""
        b = self.bufs[i % len(self.bufs)]
        return b[self.rank * self.shots : (self.rank + 1) * self.shots]

    def gathered(self, i: int):
        return self.bufs[i % len(self.bufs)]

    def mark_consumed(self, i: int, stream=None) -> None:
        """The reader of ``gathered(i)`` is done (in stream order): peers may overwrite the buffer two pushes later."""
        torch = self._torch
        ev = torch.cuda.Event()
        ev.record(stream if stream is not None else torch.cuda.current_stream())
        self._consumed[i % len(self.bufs)] = ev

    def push(self, i: int, stream=None):
        """After the producer of ``local_rows(i)`` on ``stream`` (default: current): push to all peers, barrier.
        Returns the event that marks ``gathered(i)`` complete on this rank."""
        torch = self._torch
        k = i % len(self.bufs)
        hdl = self.hdls[k]
        ready = torch.cuda.Event()
        ready.record(stream if stream is not None else torch.cuda.current_stream())
        self.side.wait_event(ready)
        # Before this rank joins barrier i, its reader of the OTHER buffer (push i - 1) must be done: peers overwrite that
        # buffer in push i + 1, which they start only after barrier i.
        prev = self._consumed[(i - 1) % len(self.bufs)] if len(self.bufs) > 1 else self._consumed[k]
        if prev is not None:
            self.side.wait_event(prev)
        src = self.bufs[k].data_ptr() + self.rank * self._slice_bytes
        with torch.cuda.stream(self.side):
            for d in range(1, self.world):
                peer = (self.rank + d) % self.world  # staggered: every link carries one transfer at a time
                dst = int(hdl.buffer_ptrs[peer]) + self.rank * self._slice_bytes
                self._check(self._lib.tsb_memcpy_peer_async(dst, src, self._slice_bytes, self.side.cuda_stream))
            hdl.barrier(channel=0)
            done = torch.cuda.Event()
            done.record(self.side)
        return done


class ShardedDetectorSampler:
    """Packed detector/observable samples from all ranks of the default process group."""

    def __init__(self, program, sparse_noise, num_f: int, *, seed: int, device: int | None = None, mode: str = "auto"):
        import torch
        import torch.distributed as dist

        nccl_defaults()  # no effect on a group that is already up; harmless then
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dp = DeviceProgram(program, device=self.device, mode=mode)
        self.noise = DeviceChannelSampler(sparse_noise, num_f, seed=seed, device=self.device)
        self._key = ((int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF)  # jax.random.key(seed)
        self._torch = torch

    def sample_packed(self, shots: int, *, batch_size: int | None = None):
        """``int64[shots, ceil(n_out/64)]`` device tensor (bit j of a row = output j), the same on every rank."""
        torch = self._torch
        if shots < 0:
            raise ValueError(f"shots must be non-negative, got {shots}")
        if batch_size is not None and batch_size < 1:
            raise ValueError(f"batch_size must be at least 1, got {batch_size}")
        wf, wo = self.dp.info["words_f64"], self.dp.info["words_out64"]
        dev = torch.device("cuda", self.device)
        if shots == 0:
            return torch.empty((0, wo), dtype=torch.int64, device=dev)
        batch = shots if batch_size is None else int(batch_size)
        stream = torch.cuda.current_stream(dev).cuda_stream
        parts = []
        for _ in range(ceil(shots / batch)):
            self._key, sub = split_key(self._key)
            call = self.noise.next_call()
            lo, hi = shard_range(batch, self.rank, self.world)
            n = hi - lo
            d_f = torch.empty((max(n, 1), wf), dtype=torch.int64, device=dev)
            d_out = torch.zeros((max(n, 1), wo), dtype=torch.int64, device=dev)
            owns_shot0 = n > 0 and lo == 0  # the normalisation check (sampler.py:66-72, 149-161) belongs to in-batch shot 0
            d_dev = torch.zeros(max(1, self.dp.info["n_components"]), dtype=torch.float32, device=dev) if owns_shot0 else None
            if n:
                self.noise.sample_device(d_f.data_ptr(), n, shot_offset=lo, call=call, stream=stream)
                self.dp.sample_device(d_f.data_ptr(), n, sub, d_out.data_ptr(), shot_offset=lo, stream=stream,
                                      d_norm_dev=d_dev.data_ptr() if owns_shot0 else 0)
            parts.append(gather_packed_rows(d_out[:n], batch, self.rank, self.world))
            if owns_shot0:
                from .sampler import check_norm_deviations

                check_norm_deviations(d_dev.cpu().numpy()[: self.dp.info["n_components"]])
        out = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
        return out[:shots]

    def sample(self, shots: int, *, batch_size: int | None = None) -> np.ndarray:
        """``bool[shots, num_outputs]`` on the host (every rank)."""
        packed = self.sample_packed(shots, batch_size=batch_size).cpu().numpy().view(np.uint64)
        n_out = self.dp.num_outputs
        if shots == 0:
            return np.zeros((0, n_out), dtype=np.bool_)
        return np.unpackbits(packed.view(np.uint8), axis=1, bitorder="little", count=n_out).astype(np.bool_)
