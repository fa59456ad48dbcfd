"""Multi-GPU plumbing: shots shard across ranks, one gather of packed output rows.

Every rank holds the whole program; rank r samples rows ``shard_range(B, r, world)`` of the batch with
``shot_offset = lo`` so the RNG counters are the in-batch shot indices and the gathered result is bit-identical
to a single-GPU run.  ``torch.distributed`` is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of ``B`` rows: sizes differ by at most one."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_bool_rows(bits: np.ndarray) -> np.ndarray:
    """``bool[B, n]`` -> ``uint64[B, ceil(n/64)]`` (the device's packed output format)."""
    B, n = bits.shape
    words = max(1, (n + 63) // 64)
    pad = words * 64 - n
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    if pad:
        b = np.concatenate([b, np.zeros((B, pad), np.uint8)], axis=1)
    return np.packbits(b, axis=1, bitorder="little").view(np.uint64).reshape(B, words)


def gather_packed_rows(rows, B: int, rank: int, world: int):
    """All-gather ragged shards of packed rows (torch int64 ``[b_r, words]``) into ``[B, words]`` on every rank."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return rows
    words = rows.shape[1]
    sizes = [shard_range(B, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros((cap, words), dtype=rows.dtype, device=rows.device)
    padded[: rows.shape[0]] = rows
    buf = torch.empty((world * cap, words), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(buf, padded)
    return torch.cat([buf[r * cap : r * cap + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
