"""Hand-built micro-programs on the APPROXIMATE branch (``has_approximate_floatfactors=True``).

The reference's rotation-gate tests (``test/integration/test_sampler_circuits.py:576-600`` R_X/R_Y/R_Z,
``:638-690`` U3 / identities / many R_X) compile circuits with stim + pyzx_param (absent here) and compare the
probabilities with closed forms.  The programs below are written directly in the ``CompiledProgram`` schema, the way
tsim's compile stage lays such circuits out: an arbitrary-angle spider ``|0> + e^{i phi pi}|1>`` is cut into two scalar
graphs, and every phase whose denominator is not in {1, 2, 4} is folded into the graph's complex64
``approximate_floatfactor`` (``src/tsim/compile/compile.py:293-303``), which switches ``evaluate`` to its float32 sum
(``src/tsim/compile/evaluate.py:56-59``).  The probability of an outcome is the doubled diagram, i.e. a product of
such two-term factors; ``expand`` multiplies the factors out into the list of graphs.

Circuit (``test_rot_gates``): ``R 0 1; H 0; CNOT 0 1; <gate> 1; [X_ERROR(p) 1;] M 0 1`` -- outcome probability
``P(m0, m1 | f) = 1/2 |<m1 ^ f| gate |m0>|^2``.
"""

from __future__ import annotations

import itertools

import numpy as np

from tsim_b200.program import CompiledComponent, make_program, make_scalar_graphs

# parameter order of the last level: [f (X error on qubit 1 before M), m0, m1]
F_BIT, M0, M1 = 0, 1, 2
N_PARAMS = 3


def _mask(*bits):
    m = np.zeros(N_PARAMS, dtype=np.uint8)
    for b in bits:
        m[b] = 1
    return m


def term(aff=1.0 + 0j, power2=0, node=(), halfpi=(), pi=(), phase_index=0):
    """One summand of a factor: ``aff * 2^power2 * w^phase_index * prod node * prod halfpi * prod pi``."""
    return dict(aff=complex(aff), power2=int(power2), node=list(node), halfpi=list(halfpi), pi=list(pi), phase_index=int(phase_index))


def phase_gadget(phi: float, mask):
    """``e^{i phi pi par(mask)} = [par = 0] + e^{i phi pi} [par = 1]``: the two cut terms of a ``Z(phi)`` spider on
    an outcome wire; ``[par = b] = (1 + w^(4 par + 4 b)) / 2`` is a node-phase term."""
    return [term(1.0, -1, node=[(0, mask)]), term(np.exp(1j * np.pi * phi), -1, node=[(4, mask)])]


def x_rotation_factor(theta: float, mask):
    """``1 + e^{i theta pi} (-1)^par(mask)``: one side of ``|<m1|R_X(theta)|m0>|^2 = (1 + (-1)^s cos(theta pi)) / 2``."""
    return [term(1.0), term(np.exp(1j * np.pi * theta), 0, halfpi=[(4, mask)])]


def expand(factors, n_params: int, extra_power2: int = 0):
    """Multiply the factors out: one scalar graph per combination of summands (``CompiledScalarGraphs``)."""
    graphs = []
    for combo in itertools.product(*factors):
        aff = np.complex128(1.0)
        g = term()
        for t in combo:
            # compile.py:295: approximate_floatfactor *= exp(1j * phase * pi) in Python complex (float64)
            aff = aff * t["aff"]
            g["power2"] += t["power2"]
            g["phase_index"] += t["phase_index"]
            g["node"] += t["node"]
            g["halfpi"] += t["halfpi"]
            g["pi"] += t["pi"]
        g["aff"] = aff
        g["power2"] += extra_power2
        graphs.append(g)
    G = len(graphs)
    A = max(1, max(len(g["node"]) for g in graphs))
    H = max(1, max(len(g["halfpi"]) for g in graphs))
    C = max(1, max(len(g["pi"]) for g in graphs))
    P = n_params
    phases, nparams, counts = np.zeros((G, A), np.uint8), np.zeros((G, A, P), np.uint8), np.zeros(G, np.int32)
    coeffs, hparams = np.zeros((G, H), np.uint8), np.zeros((G, H, P), np.uint8)
    psi_c, psi_p = np.zeros((G, C), np.uint8), np.zeros((G, C, P), np.uint8)
    phi_c, phi_p = np.zeros((G, C), np.uint8), np.zeros((G, C, P), np.uint8)
    for i, g in enumerate(graphs):
        counts[i] = len(g["node"])
        for j, (ph, m) in enumerate(g["node"]):
            phases[i, j], nparams[i, j] = ph, m[:P]
        for j, (c, m) in enumerate(g["halfpi"]):
            coeffs[i, j], hparams[i, j] = c, m[:P]
        for j, (pc, pm, fc, fm) in enumerate(g["pi"]):
            psi_c[i, j], psi_p[i, j], phi_c[i, j], phi_p[i, j] = pc, pm[:P], fc, fm[:P]
    return make_scalar_graphs(
        P,
        node=(phases, nparams, counts),
        halfpi=(coeffs, hparams),
        pi=(psi_c, psi_p, phi_c, phi_p),
        phase_indices=[g["phase_index"] % 8 for g in graphs],
        power2=[g["power2"] for g in graphs],
        approximate_floatfactors=np.array([g["aff"] for g in graphs], dtype=np.complex64),
        has_approximate_floatfactors=True,
    )


def const_approx_level(n_params: int, power2: int):
    """One graph of value ``2^power2``; marked approximate like its siblings (the flag is per level in the schema, but
    tsim sets it from the graphs of that level only -- a level without such a graph stays exact)."""
    return make_scalar_graphs(n_params, num_graphs=1, power2=[power2])


def rotation_factors(kind: str, angles):
    """Factors of ``2 P(m0, m1 | f)`` (without the common power of two) and that power."""
    s = _mask(F_BIT, M0, M1)  # parity m0 ^ m1 ^ f: the rotated qubit flips or not
    if kind == "rx":
        (theta,) = angles
        return [x_rotation_factor(theta, s), x_rotation_factor(-theta, s)], -3
    if kind == "u3":
        theta, phi, lam = angles
        # <m1'|U3|m0> = e^{i phi pi m1'} e^{i lam pi m0} (-1)^{m0 (1 - m1')} r(s), r = cos or sin of theta pi / 2;
        # m1' = m1 ^ f.  Both sides of the doubled diagram carry their gadgets; the signs square to one but stay in
        # the program as pi-product terms (psi = m0, phi = 1 + m1 + f).
        m1p = _mask(F_BIT, M1)
        sign = [term(pi=[(0, _mask(M0), 1, m1p)])]
        return (
            [
                x_rotation_factor(theta, s), x_rotation_factor(-theta, s),
                phase_gadget(phi, m1p), phase_gadget(-phi, m1p),
                phase_gadget(lam, _mask(M0)), phase_gadget(-lam, _mask(M0)),
                sign, sign,
            ],
            -3,
        )
    if kind == "many_rx":
        # n rotations by a and one by -n a, kept as separate spiders ("prevent simplification"): each side is
        # 1 + (-1)^s prod_j e^{i a_j pi}
        prod_p = np.complex128(1.0)
        prod_m = np.complex128(1.0)
        for a in angles:
            prod_p *= np.exp(1j * np.pi * a)
            prod_m *= np.exp(-1j * np.pi * a)
        return [[term(1.0), term(prod_p, halfpi=[(4, s)])], [term(1.0), term(prod_m, halfpi=[(4, s)])]], -3
    raise ValueError(kind)


def expected_probability(kind: str, angles, f: int, m0: int, m1: int) -> float:
    """Closed forms of the reference tests (``test_sampler_circuits.py:591-600`` and ``:622-635``), float64."""
    if kind == "rx":
        t = angles[0] * np.pi
        U = np.array([[np.cos(t / 2), -1j * np.sin(t / 2)], [-1j * np.sin(t / 2), np.cos(t / 2)]])
    elif kind == "u3":
        t, p, l = (a * np.pi for a in angles)
        U = np.array([[np.cos(t / 2), -np.exp(1j * l) * np.sin(t / 2)], [np.exp(1j * p) * np.sin(t / 2), np.exp(1j * (p + l)) * np.cos(t / 2)]])
    else:
        U = np.eye(2)
    return 0.5 * float(np.abs(U[m1 ^ f, m0]) ** 2)


def joint_program(kind: str, angles):
    """``CompiledStateProbs`` layout (``mode="joint"``): level 0 = normalisation (f only), level 1 = all outputs plugged."""
    factors, p2 = rotation_factors(kind, angles)
    comp = CompiledComponent((0, 1), np.array([0], np.int32), (const_approx_level(1, 0), expand(factors, 3, p2)))
    return make_program([comp], num_f=1)


def sampling_program(kind: str, angles):
    """Autoregressive layout: level 1 = P(m0 = 1) = 1/2 (m1 summed out), level 2 = P(m0, m1 = 1)."""
    factors, p2 = rotation_factors(kind, angles)
    comp = CompiledComponent(
        (0, 1), np.array([0], np.int32), (const_approx_level(1, 0), const_approx_level(2, -1), expand(factors, 3, p2))
    )
    return make_program([comp], num_f=1)


ROT_CASES = [("rx", (0.34,)), ("rx", (0.24,)), ("rx", (0.49,)), ("u3", (0.3, 0.24, 0.49)), ("u3", (0.1, -0.3, 0.2)),
             ("many_rx", (0.01, 0.01, -0.02)), ("many_rx", (0.01,) * 5 + (-0.05,))]
