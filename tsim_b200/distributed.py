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
