"""jax.random's threefry2x32 PRNG, restated in NumPy (test infrastructure only).

Third-party algorithm on the reference's path: ``jax``/``jaxlib`` 0.9.2
(pinned in the reference's ``uv.lock``); call sites in the reference:
``src/tsim/sampler.py:74-75`` (``split`` + ``bernoulli`` per output bit),
``:198`` (``jax.random.key(seed)``), ``:272,399,482`` (per-batch ``split``).

Published algorithm: Salmon et al., "Parallel random numbers: as easy as
1, 2, 3" (SC'11), Threefry-2x32 with 20 rounds, as used by ``jax.random`` with
``jax_threefry_partitionable=True`` (the default since jax 0.5): the random
word for element ``i`` of a 1-D draw is ``out0 ^ out1`` of the block cipher
applied to the 64-bit counter ``(hi, lo) = (0, i)``; ``split`` returns the
cipher outputs ``(out0, out1)`` for counters ``(0, 0)`` and ``(0, 1)``.

Pinned by the Random123 known-answer vector and by the reference's seeded
counts (``tests/test_oracle_kat.py``).
"""

from __future__ import annotations

import numpy as np

_ROT_A = (13, 15, 26, 6)
_ROT_B = (17, 29, 16, 24)
_PARITY = np.uint32(0x1BD11BDA)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32-20 block function on uint32 arrays (broadcasting)."""
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    x0 = np.array(x0, dtype=np.uint32, copy=True)
    x1 = np.array(x1, dtype=np.uint32, copy=True)
    ks = (k0, k1, k0 ^ k1 ^ _PARITY)
    with np.errstate(over="ignore"):
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for group in range(5):
            for r in _ROT_A if group % 2 == 0 else _ROT_B:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(group + 1) % 3]
            x1 = x1 + ks[(group + 2) % 3] + np.uint32(group + 1)
    return x0, x1


def key_from_seed(seed: int) -> tuple[int, int]:
    """``jax.random.key(seed)`` for a non-negative seed below 2**32 -> (0, seed)."""
    seed = int(seed)
    return (seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF


def split(key: tuple[int, int]) -> tuple[tuple[int, int], tuple[int, int]]:
    """``jax.random.split(key)`` (num=2).  Usage everywhere: ``carry, sub = split(carry)``."""
    o0, o1 = threefry2x32(key[0], key[1], np.zeros(2, np.uint32), np.arange(2, dtype=np.uint32))
    return (int(o0[0]), int(o1[0])), (int(o0[1]), int(o1[1]))


def random_bits32(key: tuple[int, int], n: int, offset: int = 0) -> np.ndarray:
    """32 random bits for elements ``offset .. offset+n-1`` of a 1-D draw."""
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    hi = (idx >> np.uint64(32)).astype(np.uint32)
    lo = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    o0, o1 = threefry2x32(key[0], key[1], hi, lo)
    return o0 ^ o1


def uniform_f32(key: tuple[int, int], n: int, offset: int = 0) -> np.ndarray:
    """``jax.random.uniform(key, (n,), float32)``: mantissa trick, [0, 1)."""
    bits = random_bits32(key, n, offset)
    fl = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32)
    return (fl - np.float32(1.0)).astype(np.float32)
