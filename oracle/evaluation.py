"""``evaluate`` and the four term families, restated in NumPy (test infrastructure only).

Follows reference ``src/tsim/compile/terms.py:22-207`` (tables and the
``evaluate`` method of each family), ``src/tsim/utils/linalg.py:81-102``
(``matmul_gf2`` as a float32 GEMM followed by ``% 2``) and
``src/tsim/compile/evaluate.py:15-59`` (six-way product, exact and approximate
branch).  ``circuit`` is a ``tsim_b200.program.CompiledScalarGraphs`` or any
object with the reference's field names.

float32 tail fixed by this oracle (the reference leaves it to XLA; unpinned):

* ``to_complex``: see ``oracle.exact_scalar.to_complex_parts``.
* ``complex_abs``: XLA's published lowering of ``abs(complex64)``
  (``EmitComplexAbs``): ``mx = max(|re|,|im|); mn = min(|re|,|im|);
  r = mx * sqrt(1 + (mn/mx)**2)``; ``mn`` where ``r`` is NaN.  float32, no FMA.
* approximate branch: per graph ``t = to_complex(T_g)``;
  ``u = (t.re*a.re - t.im*a.im, t.re*a.im + t.im*a.re)``; ``v = u * 2**power2``;
  the sum over graphs runs sequentially ``g = 0 .. G-1`` in float32.
"""

from __future__ import annotations

import numpy as np

from .exact_scalar import ExactScalar, pow2_f32, to_complex_parts

UNIT_PHASES = np.array(
    [
        [1, 0, 0, 0],
        [0, 1, 0, 0],
        [0, 0, 1, 0],
        [0, 0, 0, -1],
        [-1, 0, 0, 0],
        [0, -1, 0, 0],
        [0, 0, -1, 0],
        [0, 0, 0, 1],
    ],
    dtype=np.int32,
)
ONE_PLUS_PHASES = UNIT_PHASES.copy()
ONE_PLUS_PHASES[:, 0] += 1
IDENTITY = np.array([1, 0, 0, 0], dtype=np.int32)


def matmul_gf2(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """``a[G,T,P] x b[B,P] -> u8[B,G,T]`` parities, via float32 GEMM then ``% 2``."""
    G, T, _ = a.shape
    if G * T == 0:
        return np.zeros((b.shape[0], G, T), dtype=np.uint8)
    s = b.astype(np.float32) @ a.astype(np.float32).reshape(G * T, -1).T
    return (s.reshape(-1, G, T) % 2).astype(np.uint8)


def node_phases(np_, x) -> ExactScalar:
    rowsum = matmul_gf2(np_.params, x).astype(np.int32)
    idx = (4 * rowsum + np_.phases.astype(np.int32)) % 8
    vals = ONE_PLUS_PHASES[idx]
    mask = np.arange(np_.phases.shape[1])[None, :] < np_.counts[:, None]
    vals = np.where(mask[None, :, :, None], vals, IDENTITY)
    return ExactScalar.of(vals).prod(axis=-1)


def halfpi_phases(hp, x) -> ExactScalar:
    rowsum = matmul_gf2(hp.params, x).astype(np.int32)
    idx = (rowsum * hp.coeffs.astype(np.int32)) % 8
    total = np.sum(idx, axis=-1) % 8
    return ExactScalar.of(UNIT_PHASES[total])


def pi_products(pp, x) -> ExactScalar:
    psi = (pp.psi_const.astype(np.int32) + matmul_gf2(pp.psi_params, x)) % 2
    phi = (pp.phi_const.astype(np.int32) + matmul_gf2(pp.phi_params, x)) % 2
    e = np.sum((psi * phi) % 2, axis=-1) % 2
    return ExactScalar.of((1 - 2 * e)[..., None].astype(np.int32) * IDENTITY)


def phase_pairs(pr, x) -> ExactScalar:
    ra = matmul_gf2(pr.alpha_params, x).astype(np.int32)
    rb = matmul_gf2(pr.beta_params, x).astype(np.int32)
    alpha = (pr.alpha.astype(np.int32) + 4 * ra) % 8
    beta = (pr.beta.astype(np.int32) + 4 * rb) % 8
    gamma = (alpha + beta) % 8
    vals = IDENTITY + UNIT_PHASES[alpha] + UNIT_PHASES[beta] - UNIT_PHASES[gamma]
    mask = np.arange(pr.alpha.shape[1])[None, :] < pr.counts[:, None]
    vals = np.where(mask[None, :, :, None], vals, IDENTITY).astype(np.int32)
    return ExactScalar.of(vals).prod(axis=-1)


def term_product(circuit, x) -> ExactScalar:
    """``T[b,g]``: the six-way plain product of ``evaluate.py:40-50`` (before any sum)."""
    pre = circuit.prefactor
    B = x.shape[0]
    G = pre.phase_indices.shape[0]
    static = ExactScalar.of(np.broadcast_to(UNIT_PHASES[pre.phase_indices], (B, G, 4)))
    ff = ExactScalar.of(np.broadcast_to(pre.floatfactor.astype(np.int32), (B, G, 4)))
    total = node_phases(circuit.node_phases, x)
    for fac in (
        halfpi_phases(circuit.halfpi_phases, x),
        pi_products(circuit.pi_products, x),
        phase_pairs(circuit.phase_pairs, x),
        static,
        ff,
    ):
        total = total * fac
    return total


def evaluate_parts(circuit, x):
    """(re, im) float32 arrays of shape ``[B]``."""
    x = np.asarray(x).astype(np.uint8)
    pre = circuit.prefactor
    B = x.shape[0]
    if pre.phase_indices.shape[0] == 0:
        z = np.zeros(B, dtype=np.float32)
        return z, z.copy()
    total = term_product(circuit, x)
    if not pre.has_approximate_floatfactors:
        with np.errstate(over="ignore"):
            s = ExactScalar(total.coeffs, (total.power + pre.power2[None, :]).astype(np.int32)).sum()
        return to_complex_parts(s.coeffs, s.power)
    tre, tim = to_complex_parts(total.coeffs, total.power)  # [B, G]
    are = pre.approximate_floatfactors.real.astype(np.float32)
    aim = pre.approximate_floatfactors.imag.astype(np.float32)
    pw = pow2_f32(pre.power2)
    acc_re = np.zeros(B, dtype=np.float32)
    acc_im = np.zeros(B, dtype=np.float32)
    with np.errstate(all="ignore"):
        for g in range(tre.shape[1]):
            ure = ((tre[:, g] * are[g]).astype(np.float32) - (tim[:, g] * aim[g]).astype(np.float32)).astype(np.float32)
            uim = ((tre[:, g] * aim[g]).astype(np.float32) + (tim[:, g] * are[g]).astype(np.float32)).astype(np.float32)
            acc_re = (acc_re + (ure * pw[g]).astype(np.float32)).astype(np.float32)
            acc_im = (acc_im + (uim * pw[g]).astype(np.float32)).astype(np.float32)
    return acc_re, acc_im


def evaluate(circuit, x) -> np.ndarray:
    """Amplitude per row of ``x``; complex64 ``[B]`` (reference ``evaluate``)."""
    re, im = evaluate_parts(circuit, x)
    out = np.empty(re.shape, dtype=np.complex64)
    out.real = re
    out.imag = im
    return out


def complex_abs(re: np.ndarray, im: np.ndarray) -> np.ndarray:
    """float32 ``|re + i*im|`` with XLA's max*sqrt(1+(min/max)^2) lowering."""
    a = np.abs(re.astype(np.float32))
    b = np.abs(im.astype(np.float32))
    with np.errstate(all="ignore"):
        # XLA's max/min propagate NaN; a NaN ends in the NaN select below either way.
        mx = np.maximum(a, b)
        mn = np.minimum(a, b)
        r = (mn / mx).astype(np.float32)
        t = (np.float32(1.0) + (r * r).astype(np.float32)).astype(np.float32)
        res = (mx * np.sqrt(t).astype(np.float32)).astype(np.float32)
    return np.where(np.isnan(res), mn, res).astype(np.float32)


def evaluate_abs(circuit, x) -> np.ndarray:
    re, im = evaluate_parts(circuit, x)
    return complex_abs(re, im)
