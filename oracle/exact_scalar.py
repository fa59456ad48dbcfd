"""Exact Z[w]*2^p arithmetic of the reference, restated in NumPy (test infrastructure only).

Follows reference ``src/tsim/core/exact_scalar.py``:

* ``mul``              -- ``_scalar_mul`` (:19-39)
* ``reduce_step``      -- ``_reduce_power_coeffs_step`` (:42-49)
* ``mul_with_power``   -- ``_scalar_mul_with_power`` (:52-71)
* ``add_with_power``   -- ``_scalar_add_with_power`` (:74-84)
* ``fold``             -- ``_reduce_along_scan`` (:98-137): sequential carry, one
  ``reduce_step`` per combine, then reduce to a fixpoint
* ``to_complex``       -- ``_scalar_to_complex`` (:87-89) + ``to_complex`` (:218-222)

A scalar is ``(c0 + c1*w + c2*i + c3*conj(w)) * 2**p`` with ``w = exp(i*pi/4)``;
coefficients and powers are int32 with two's-complement wrap-around, exactly
like the reference's jnp.int32 arrays.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# float32(cos(pi/4)) == float32(sin(pi/4)) == 0x3F3504F3 ("f32 tail")
SQRT1_2_F32 = np.array([0x3F3504F3], dtype=np.uint32).view(np.float32)[0]


@dataclass
class ExactScalar:
    coeffs: np.ndarray  # int32 [..., 4]
    power: np.ndarray  # int32 [...]

    @staticmethod
    def of(coeffs, power=None) -> "ExactScalar":
        coeffs = np.asarray(coeffs, dtype=np.int32)
        if power is None:
            power = np.zeros(coeffs.shape[:-1], dtype=np.int32)
        return ExactScalar(coeffs, np.asarray(power, dtype=np.int32))

    def __mul__(self, other: "ExactScalar") -> "ExactScalar":
        """Plain product, no reduction (reference ``__mul__``, :167-171)."""
        with np.errstate(over="ignore"):
            return ExactScalar(mul(self.coeffs, other.coeffs), (self.power + other.power).astype(np.int32))

    def prod(self, axis: int = -1) -> "ExactScalar":
        if axis < 0:
            axis += self.power.ndim
        if self.coeffs.shape[axis] == 0:
            shape = self.coeffs.shape[:axis] + self.coeffs.shape[axis + 1 :]
            c = np.zeros(shape, dtype=np.int32)
            c[..., 0] = 1
            return ExactScalar.of(c)
        p, c = fold(self.power, self.coeffs, mul_with_power, axis)
        return ExactScalar(c, p)

    def sum(self, axis: int = -1) -> "ExactScalar":
        if axis < 0:
            axis += self.power.ndim
        p, c = fold(self.power, self.coeffs, add_with_power, axis)
        return ExactScalar(c, p)

    def to_complex(self) -> np.ndarray:
        return to_complex(self.coeffs, self.power)


def mul(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    a1, b1, c1, d1 = (x[..., i] for i in range(4))
    a2, b2, c2, d2 = (y[..., i] for i in range(4))
    with np.errstate(over="ignore"):
        A = a1 * a2 + b1 * d2 - c1 * c2 + d1 * b2
        B = a1 * b2 + b1 * a2 + c1 * d2 + d1 * c2
        C = a1 * c2 + b1 * b2 + c1 * a2 - d1 * d2
        D = a1 * d2 - b1 * c2 - c1 * b2 + d1 * a2
    return np.stack([A, B, C, D], axis=-1).astype(np.int32)


def reduce_step(power: np.ndarray, coeffs: np.ndarray):
    reducible = np.all((coeffs & 1) == 0, axis=-1) & np.any(coeffs != 0, axis=-1)
    coeffs = np.where(reducible[..., None], coeffs >> 1, coeffs).astype(np.int32)
    power = np.where(reducible, power + 1, power).astype(np.int32)
    return power, coeffs


def mul_with_power(x, y):
    p1, c1 = x
    p2, c2 = y
    with np.errstate(over="ignore"):
        return reduce_step((p1 + p2).astype(np.int32), mul(c1, c2))


def _shl_one(n) -> np.ndarray:
    """int32 ``1 << n`` with XLA semantics: shift counts >= 32 give 0."""
    n = np.asarray(n, dtype=np.int64)
    v = np.where(n >= 32, 0, np.left_shift(np.int64(1), np.minimum(n, 31)))
    return np.asarray(v, dtype=np.int64).astype(np.uint32).astype(np.int32)


def add_with_power(x, y):
    p1, c1 = x
    p2, c2 = y
    p1 = np.asarray(p1, dtype=np.int32)
    p2 = np.asarray(p2, dtype=np.int32)
    with np.errstate(over="ignore"):
        s1 = _shl_one(np.maximum(p1.astype(np.int64) - p2, 0))
        s2 = _shl_one(np.maximum(p2.astype(np.int64) - p1, 0))
        p = np.minimum(p1, p2).astype(np.int32)
        c = (c1 * s1[..., None] + c2 * s2[..., None]).astype(np.int32)
    return reduce_step(p, c)


def fold(power: np.ndarray, coeffs: np.ndarray, op, axis: int):
    if axis < 0:
        axis += power.ndim
    pt = np.moveaxis(power, axis, 0)
    ct = np.moveaxis(coeffs, axis, 0)
    carry = (pt[0].astype(np.int32), ct[0].astype(np.int32))
    for i in range(1, pt.shape[0]):
        carry = op(carry, (pt[i], ct[i]))
    p, c = carry
    while True:
        new_p, new_c = reduce_step(p, c)
        changed = bool(np.any(new_p != p))
        p, c = new_p, new_c
        if not changed:
            break
    return p, c


def pow2_f32(p: np.ndarray) -> np.ndarray:
    """Exact float32 ``2.0**p`` (denormals kept, overflow -> inf)."""
    p = np.asarray(p, dtype=np.int64)
    bits = np.where(
        p > 127,
        0x7F800000,
        np.where(p >= -126, (np.clip(p, -126, 127) + 127) << 23, np.where(p >= -149, 1 << np.clip(p + 149, 0, 22), 0)),
    )
    return bits.astype(np.uint32).view(np.float32)


def to_complex_parts(coeffs: np.ndarray, power: np.ndarray):
    """(re, im) float32; fixed op order, one rounding per op, no FMA.

    ``re = ((f(c0) + f(c1)*s) + 0) + f(c3)*s``; ``im = ((0 + f(c1)*s) + f(c2)) - f(c3)*s``;
    then each part times ``2**p``.
    """
    s = SQRT1_2_F32
    f = coeffs.astype(np.float32)
    c0, c1, c2, c3 = (f[..., i] for i in range(4))
    with np.errstate(all="ignore"):
        t1 = (c1 * s).astype(np.float32)
        t3 = (c3 * s).astype(np.float32)
        re = ((c0 + t1).astype(np.float32) + t3).astype(np.float32)
        im = ((t1 + c2).astype(np.float32) - t3).astype(np.float32)
        sc = pow2_f32(power)
        return (re * sc).astype(np.float32), (im * sc).astype(np.float32)


def to_complex(coeffs: np.ndarray, power: np.ndarray) -> np.ndarray:
    re, im = to_complex_parts(coeffs, power)
    out = np.empty(re.shape, dtype=np.complex64)
    out.real = re
    out.imag = im
    return out
