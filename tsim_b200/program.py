"""Host-side program schema for the sampling hot path.

These are plain-NumPy mirrors of the reference's containers -- same field
names, same shapes, same dtypes -- so that a ``CompiledProgram`` produced by
tsim's (unchanged) compile stages can be handed over attribute by attribute:

* ``NodePhases``/``HalfPiPhases``/``PiProducts``/``PhasePairs``/``ScalarPrefactor``
  -- reference ``src/tsim/compile/terms.py:42-207``
* ``CompiledScalarGraphs`` -- reference ``src/tsim/compile/compile.py:21-37``
* ``CompiledComponent`` / ``CompiledProgram`` -- reference
  ``src/tsim/core/types.py:55-107``

``from_tsim`` converts real tsim objects (jax arrays) by duck typing;
``save_npz``/``load_npz`` give the travel format (no jax/equinox needed to
load).  ``pack_program`` turns a program into the flat, bit-packed blob the
C-ABI uploads to HBM (layout documented in DESIGN.md and
``include/tsim_b200.h``).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np

# ----------------------------------------------------------------------------
# Containers (reference field names)
# ----------------------------------------------------------------------------


def _u8(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(np.asarray(a), dtype=np.uint8)
    if shape is not None:
        out = out.reshape(shape)
    return out


def _i32(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(np.asarray(a), dtype=np.int32)
    if shape is not None:
        out = out.reshape(shape)
    return out


@dataclass
class NodePhases:
    """``prod_j (1 + w^(4*parity_j + phases_j))``; reference terms.py:42-73."""

    phases: np.ndarray  # u8 [G, A] values 0..7
    params: np.ndarray  # u8 [G, A, P] 0/1
    counts: np.ndarray  # i32 [G]


@dataclass
class HalfPiPhases:
    """``w^(sum_j coeffs_j * parity_j)``; reference terms.py:76-107."""

    coeffs: np.ndarray  # u8 [G, H] in {0,2,4,6}
    params: np.ndarray  # u8 [G, H, P]


@dataclass
class PiProducts:
    """``(-1)^(sum_j psi_j*phi_j)``; reference terms.py:110-144."""

    psi_const: np.ndarray  # u8 [G, C]
    psi_params: np.ndarray  # u8 [G, C, P]
    phi_const: np.ndarray  # u8 [G, C]
    phi_params: np.ndarray  # u8 [G, C, P]


@dataclass
class PhasePairs:
    """``prod_j (1 + w^a + w^b - w^(a+b))``; reference terms.py:147-187."""

    alpha: np.ndarray  # u8 [G, D]
    alpha_params: np.ndarray  # u8 [G, D, P]
    beta: np.ndarray  # u8 [G, D]
    beta_params: np.ndarray  # u8 [G, D, P]
    counts: np.ndarray  # i32 [G]


@dataclass
class ScalarPrefactor:
    """Static per-graph prefactor; reference terms.py:190-207."""

    phase_indices: np.ndarray  # u8 [G]
    floatfactor: np.ndarray  # i32 [G, 4]
    power2: np.ndarray  # i32 [G]
    approximate_floatfactors: np.ndarray  # c64 [G]
    has_approximate_floatfactors: bool = False


@dataclass
class CompiledScalarGraphs:
    """One level of a component: ``G`` scalar graphs over ``P`` parameters."""

    num_graphs: int
    n_params: int
    node_phases: NodePhases
    halfpi_phases: HalfPiPhases
    pi_products: PiProducts
    phase_pairs: PhasePairs
    prefactor: ScalarPrefactor


@dataclass
class CompiledComponent:
    """Reference core/types.py:55-77."""

    output_indices: tuple[int, ...]
    f_selection: np.ndarray  # i32 [F_c]
    compiled_scalar_graphs: tuple[CompiledScalarGraphs, ...]


@dataclass
class CompiledProgram:
    """Reference core/types.py:80-107."""

    components: tuple[CompiledComponent, ...]
    direct_f_indices: np.ndarray  # i32 [n_direct]
    direct_flips: np.ndarray  # bool [n_direct]
    output_order: np.ndarray  # i32 [n_out]
    output_reindex: np.ndarray | None  # i32 [n_out] or None
    num_outputs: int
    num_detectors: int
    # Not in the reference container: width of the global f vector.  The
    # reference reads it off ``f_params.shape[1]``; the device needs it at
    # upload time to size the bit-packed f rows.  ``None`` -> inferred.
    num_f: int | None = None
    meta: dict = field(default_factory=dict)

    def infer_num_f(self) -> int:
        if self.num_f is not None:
            return int(self.num_f)
        hi = -1
        if len(self.direct_f_indices):
            hi = max(hi, int(np.max(self.direct_f_indices)))
        for c in self.components:
            if len(c.f_selection):
                hi = max(hi, int(np.max(c.f_selection)))
        return hi + 1


# ----------------------------------------------------------------------------
# Builders
# ----------------------------------------------------------------------------


def make_scalar_graphs(
    n_params: int,
    *,
    node=None,
    halfpi=None,
    pi=None,
    pairs=None,
    phase_indices=None,
    floatfactor=None,
    power2=None,
    approximate_floatfactors=None,
    has_approximate_floatfactors: bool | None = None,
    num_graphs: int | None = None,
) -> CompiledScalarGraphs:
    """Assemble a level from optional families, padding the missing ones.

    ``node=(phases[G,A], params[G,A,P], counts[G])``, ``halfpi=(coeffs, params)``,
    ``pi=(psi_const, psi_params, phi_const, phi_params)``,
    ``pairs=(alpha, alpha_params, beta, beta_params, counts)``.
    """
    P = int(n_params)
    if num_graphs is None:
        for fam in (node, halfpi, pi, pairs):
            if fam is not None:
                num_graphs = int(np.asarray(fam[0]).shape[0])
                break
        else:
            num_graphs = 0 if phase_indices is None else len(phase_indices)
    G = int(num_graphs)

    def _fam_params(a, T):
        return _u8(a).reshape(G, T, P)

    def _width(a) -> int:
        a = np.asarray(a)
        if a.ndim >= 2:
            return int(a.shape[1])
        return int(a.size // G) if G else 0

    if node is None:
        node = (np.zeros((G, 0)), np.zeros((G, 0, P)), np.zeros((G,)))
    A = _width(node[0])
    np_ = NodePhases(_u8(node[0], (G, A)), _fam_params(node[1], A), _i32(node[2], (G,)))

    if halfpi is None:
        halfpi = (np.zeros((G, 0)), np.zeros((G, 0, P)))
    H = _width(halfpi[0])
    hp = HalfPiPhases(_u8(halfpi[0], (G, H)), _fam_params(halfpi[1], H))

    if pi is None:
        pi = (np.zeros((G, 0)), np.zeros((G, 0, P)), np.zeros((G, 0)), np.zeros((G, 0, P)))
    C = _width(pi[0])
    pp = PiProducts(
        _u8(pi[0], (G, C)), _fam_params(pi[1], C), _u8(pi[2], (G, C)), _fam_params(pi[3], C)
    )

    if pairs is None:
        pairs = (
            np.zeros((G, 0)),
            np.zeros((G, 0, P)),
            np.zeros((G, 0)),
            np.zeros((G, 0, P)),
            np.zeros((G,)),
        )
    D = _width(pairs[0])
    pr = PhasePairs(
        _u8(pairs[0], (G, D)),
        _fam_params(pairs[1], D),
        _u8(pairs[2], (G, D)),
        _fam_params(pairs[3], D),
        _i32(pairs[4], (G,)),
    )

    if phase_indices is None:
        phase_indices = np.zeros(G)
    if floatfactor is None:
        floatfactor = np.tile(np.array([1, 0, 0, 0]), (G, 1))
    if power2 is None:
        power2 = np.zeros(G)
    if approximate_floatfactors is None:
        approximate_floatfactors = np.ones(G, dtype=np.complex64)
    aff = np.ascontiguousarray(np.asarray(approximate_floatfactors), dtype=np.complex64).reshape(G)
    if has_approximate_floatfactors is None:
        # reference compile.py:300-302: any(aff != 1.0)
        has_approximate_floatfactors = bool(np.any(aff != np.complex64(1.0)))
    pre = ScalarPrefactor(
        _u8(phase_indices, (G,)),
        _i32(floatfactor, (G, 4)),
        _i32(power2, (G,)),
        aff,
        bool(has_approximate_floatfactors),
    )
    return CompiledScalarGraphs(G, P, np_, hp, pp, pr, pre)


def empty_scalar_graphs(n_params: int) -> CompiledScalarGraphs:
    """Level with zero graphs: ``evaluate`` returns 0 (reference evaluate.py:34-35)."""
    return make_scalar_graphs(n_params, num_graphs=0)


def make_program(
    components,
    *,
    direct_f_indices=(),
    direct_flips=(),
    output_order=None,
    num_outputs: int | None = None,
    num_detectors: int | None = None,
    num_f: int | None = None,
    meta: dict | None = None,
) -> CompiledProgram:
    """Assemble a program; ``output_reindex`` is derived as in pipeline.py:90-99."""
    components = tuple(components)
    direct_f_indices = _i32(direct_f_indices).reshape(-1)
    direct_flips = np.asarray(direct_flips, dtype=np.bool_).reshape(-1)
    n_compiled = sum(len(c.output_indices) for c in components)
    n_out = len(direct_f_indices) + n_compiled
    if output_order is None:
        output_order = np.arange(n_out, dtype=np.int32)
    output_order = _i32(output_order).reshape(-1)
    if num_outputs is None:
        num_outputs = n_out
    reindex = np.argsort(output_order, kind="stable").astype(np.int32)
    is_identity = np.array_equal(reindex, np.arange(len(output_order)))
    prog = CompiledProgram(
        components=components,
        direct_f_indices=direct_f_indices,
        direct_flips=direct_flips,
        output_order=output_order,
        output_reindex=None if is_identity else reindex,
        num_outputs=int(num_outputs),
        num_detectors=int(num_outputs if num_detectors is None else num_detectors),
        num_f=num_f,
        meta=dict(meta or {}),
    )
    return prog


# ----------------------------------------------------------------------------
# tsim adapter (duck typed; works on jax arrays / equinox modules)
# ----------------------------------------------------------------------------


def _np(a) -> np.ndarray:
    return np.asarray(a)


def scalar_graphs_from_tsim(csg: Any) -> CompiledScalarGraphs:
    """Convert a tsim ``CompiledScalarGraphs`` (compile.py:21-37) to NumPy."""
    G = int(csg.num_graphs)
    P = int(csg.n_params)
    n, h, p, q, f = csg.node_phases, csg.halfpi_phases, csg.pi_products, csg.phase_pairs, csg.prefactor
    return make_scalar_graphs(
        P,
        num_graphs=G,
        node=(_np(n.phases), _np(n.params), _np(n.counts)),
        halfpi=(_np(h.coeffs), _np(h.params)),
        pi=(_np(p.psi_const), _np(p.psi_params), _np(p.phi_const), _np(p.phi_params)),
        pairs=(_np(q.alpha), _np(q.alpha_params), _np(q.beta), _np(q.beta_params), _np(q.counts)),
        phase_indices=_np(f.phase_indices),
        floatfactor=_np(f.floatfactor),
        power2=_np(f.power2),
        approximate_floatfactors=_np(f.approximate_floatfactors),
        has_approximate_floatfactors=bool(f.has_approximate_floatfactors),
    )


def from_tsim(program: Any, *, num_f: int | None = None) -> CompiledProgram:
    """Convert a tsim ``CompiledProgram`` (core/types.py:80-107) to NumPy."""
    if isinstance(program, CompiledProgram):
        if num_f is not None and int(num_f) > program.infer_num_f():
            from dataclasses import replace

            return replace(program, num_f=int(num_f))  # wider f rows than the program references: columns ignored
        return program
    comps = []
    for c in program.components:
        comps.append(
            CompiledComponent(
                output_indices=tuple(int(i) for i in c.output_indices),
                f_selection=_i32(_np(c.f_selection)).reshape(-1),
                compiled_scalar_graphs=tuple(
                    scalar_graphs_from_tsim(g) for g in c.compiled_scalar_graphs
                ),
            )
        )
    reindex = program.output_reindex
    return CompiledProgram(
        components=tuple(comps),
        direct_f_indices=_i32(_np(program.direct_f_indices)).reshape(-1),
        direct_flips=np.asarray(_np(program.direct_flips), dtype=np.bool_).reshape(-1),
        output_order=_i32(_np(program.output_order)).reshape(-1),
        output_reindex=None if reindex is None else _i32(_np(reindex)).reshape(-1),
        num_outputs=int(program.num_outputs),
        num_detectors=int(program.num_detectors),
        num_f=num_f,
    )


# ----------------------------------------------------------------------------
# .npz travel format
# ----------------------------------------------------------------------------

_LEVEL_FIELDS = (
    ("node_phases", ("phases", "params", "counts")),
    ("halfpi_phases", ("coeffs", "params")),
    ("pi_products", ("psi_const", "psi_params", "phi_const", "phi_params")),
    ("phase_pairs", ("alpha", "alpha_params", "beta", "beta_params", "counts")),
    ("prefactor", ("phase_indices", "floatfactor", "power2", "approximate_floatfactors")),
)


def save_npz(path: str, program: CompiledProgram, *, noise: Any = None, meta: dict | None = None) -> None:
    """Write a program as a flat ``.npz`` (mask tensors bit-packed along P).

    ``noise``: optional channel sampler (tsim's or ours) or its ``_sparse_data`` list of ``(p_fire, cond_cdf, xor_patterns)``
    (reference ``noise/channels.py:578-622``); stored so that a benchmark can draw f vectors from the circuit's own noise
    model.  ``meta``: free-form strings (circuit name, tsim version, ...)."""
    d: dict[str, np.ndarray] = {}
    if noise is not None:
        sparse = getattr(noise, "_sparse_data", noise)
        d["noise.n"] = np.array([len(sparse)], dtype=np.int64)
        d["noise.p_fire"] = np.array([float(t[0]) for t in sparse], dtype=np.float64)
        d["noise.cdf_len"] = np.array([len(t[1]) for t in sparse], dtype=np.int64)
        d["noise.cdf"] = np.concatenate([np.asarray(t[1], np.float64).reshape(-1) for t in sparse]) if sparse else np.zeros(0)
        pats = [np.asarray(t[2], np.uint8).reshape(len(t[1]), -1) for t in sparse]
        d["noise.num_f"] = np.array([pats[0].shape[1] if pats else (program.num_f or 0)], dtype=np.int64)
        d["noise.patterns"] = (
            np.packbits(np.concatenate(pats, axis=0), axis=1, bitorder="little") if pats else np.zeros((0, 0), np.uint8)
        )
    for k, v in (meta or {}).items():
        d["meta." + str(k)] = np.array([str(v)])
    d["header"] = np.array(
        [
            program.num_outputs,
            program.num_detectors,
            len(program.components),
            -1 if program.num_f is None else program.num_f,
            0 if program.output_reindex is None else 1,
        ],
        dtype=np.int64,
    )
    d["direct_f_indices"] = program.direct_f_indices
    d["direct_flips"] = program.direct_flips.astype(np.uint8)
    d["output_order"] = program.output_order
    if program.output_reindex is not None:
        d["output_reindex"] = program.output_reindex
    for ci, c in enumerate(program.components):
        d[f"c{ci}.output_indices"] = np.asarray(c.output_indices, dtype=np.int32)
        d[f"c{ci}.f_selection"] = c.f_selection
        d[f"c{ci}.n_levels"] = np.array([len(c.compiled_scalar_graphs)], dtype=np.int64)
        for li, lv in enumerate(c.compiled_scalar_graphs):
            pre = f"c{ci}.l{li}."
            d[pre + "shape"] = np.array(
                [lv.num_graphs, lv.n_params, int(lv.prefactor.has_approximate_floatfactors)],
                dtype=np.int64,
            )
            for fam, names in _LEVEL_FIELDS:
                obj = getattr(lv, fam)
                for nm in names:
                    arr = getattr(obj, nm)
                    if nm.endswith("params"):
                        d[pre + fam + "." + nm + ".shape"] = np.array(arr.shape, dtype=np.int64)
                        arr = np.packbits(arr.astype(np.uint8), axis=-1, bitorder="little")
                    d[pre + fam + "." + nm] = arr
    np.savez_compressed(path, **d)


def load_npz_noise(path: str):
    """The noise tables stored by :func:`save_npz` as ``(sparse_data, num_f)``, or ``None``."""
    z = np.load(path)
    if "noise.n" not in z.files:
        return None
    num_f = int(z["noise.num_f"][0])
    lens = [int(v) for v in z["noise.cdf_len"]]
    cdf, pats = z["noise.cdf"], z["noise.patterns"]
    pats = np.unpackbits(pats, axis=1, bitorder="little", count=num_f) if pats.size else np.zeros((0, num_f), np.uint8)
    out, pos = [], 0
    for p_fire, n in zip(z["noise.p_fire"], lens):
        out.append((float(p_fire), np.array(cdf[pos : pos + n], dtype=np.float64), np.ascontiguousarray(pats[pos : pos + n])))
        pos += n
    return out, num_f


def load_npz_meta(path: str) -> dict:
    z = np.load(path)
    return {k[5:]: str(z[k][0]) for k in z.files if k.startswith("meta.")}


def load_npz(path: str) -> CompiledProgram:
    """Inverse of :func:`save_npz` (the program part)."""
    z = np.load(path)
    n_out, n_det, n_comp, num_f, has_reindex = (int(v) for v in z["header"])
    comps = []
    for ci in range(n_comp):
        n_levels = int(z[f"c{ci}.n_levels"][0])
        levels = []
        for li in range(n_levels):
            pre = f"c{ci}.l{li}."
            G, P, has_aff = (int(v) for v in z[pre + "shape"])
            vals: dict[str, dict[str, np.ndarray]] = {}
            for fam, names in _LEVEL_FIELDS:
                vals[fam] = {}
                for nm in names:
                    arr = z[pre + fam + "." + nm]
                    if nm.endswith("params"):
                        shp = tuple(int(v) for v in z[pre + fam + "." + nm + ".shape"])
                        arr = np.unpackbits(arr, axis=-1, bitorder="little", count=shp[-1]) if shp[-1] else np.zeros(shp, np.uint8)
                        arr = arr.reshape(shp)
                    vals[fam][nm] = arr
            n, h, p, q, f = (vals[k] for k, _ in _LEVEL_FIELDS)
            levels.append(
                make_scalar_graphs(
                    P,
                    num_graphs=G,
                    node=(n["phases"], n["params"], n["counts"]),
                    halfpi=(h["coeffs"], h["params"]),
                    pi=(p["psi_const"], p["psi_params"], p["phi_const"], p["phi_params"]),
                    pairs=(q["alpha"], q["alpha_params"], q["beta"], q["beta_params"], q["counts"]),
                    phase_indices=f["phase_indices"],
                    floatfactor=f["floatfactor"],
                    power2=f["power2"],
                    approximate_floatfactors=f["approximate_floatfactors"],
                    has_approximate_floatfactors=bool(has_aff),
                )
            )
        comps.append(
            CompiledComponent(
                output_indices=tuple(int(v) for v in z[f"c{ci}.output_indices"]),
                f_selection=_i32(z[f"c{ci}.f_selection"]).reshape(-1),
                compiled_scalar_graphs=tuple(levels),
            )
        )
    return CompiledProgram(
        components=tuple(comps),
        direct_f_indices=_i32(z["direct_f_indices"]).reshape(-1),
        direct_flips=z["direct_flips"].astype(np.bool_).reshape(-1),
        output_order=_i32(z["output_order"]).reshape(-1),
        output_reindex=_i32(z["output_reindex"]).reshape(-1) if has_reindex else None,
        num_outputs=n_out,
        num_detectors=n_det,
        num_f=None if num_f < 0 else num_f,
    )


# ----------------------------------------------------------------------------
# Statistics (reference sampler.py:557-609 __repr__ numbers)
# ----------------------------------------------------------------------------


def program_stats(program: CompiledProgram) -> dict:
    """Compile statistics in the vocabulary of the reference's ``__repr__``."""
    graphs = a = b = c = d = 0
    max_params = 0
    max_outputs = 0
    for comp in program.components:
        max_outputs = max(max_outputs, len(comp.output_indices))
        for lv in comp.compiled_scalar_graphs:
            graphs += lv.num_graphs
            max_params = max(max_params, lv.n_params)
            a += lv.node_phases.phases.size
            b += lv.halfpi_phases.coeffs.size
            c += lv.pi_products.psi_const.size
            d += lv.phase_pairs.alpha.size + lv.phase_pairs.beta.size
    return {
        "direct": int(len(program.direct_f_indices)),
        "graphs": int(graphs),
        "max_outputs_per_component": int(max_outputs),
        "max_params": int(max_params),
        "A_terms": int(a),
        "B_terms": int(b),
        "C_terms": int(c),
        "D_terms": int(d),
    }
