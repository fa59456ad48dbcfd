"""Shape-matched synthetic programs for the benchmark configurations.

tsim's compile stages (stim -> ZX -> stabiliser-rank decomposition) need ``stim`` and ``pyzx_param``,
which are not available in the build or benchmark environment, so the real g_{tki} of the named
circuits cannot be produced there.  These generators emit ``CompiledProgram`` objects with the
*shapes* of those programs (number of graphs per level from the published
stabiliser-term counts, parameter counts from the circuit structure, padded term counts per family)
and i.i.d. random contents, so the device does the same amount and kind of work per shot.  Real
programs dumped on a machine that has tsim (``program.from_tsim`` + ``program.save_npz``) drop into
the same code path.

The marginals of a random program do not form a probability tree; each level's ``power2`` is
shifted by a common integer so that ``p1/prev`` is of order one half for the noiseless shot, which
keeps the drawn bits non-degenerate.  Norm-deviation warnings are meaningless for these programs.
"""

from __future__ import annotations

import numpy as np

from .program import CompiledComponent, CompiledProgram, CompiledScalarGraphs, make_program, make_scalar_graphs

CONFIGS = {
    # name: dict(n_direct, components=[(n_c, F_c, G per level)], num_f, A, H, C, D, approx)
    "cfg2_distill35": dict(
        n_direct=15, comps=[(5, 48, (16, 20, 24, 28, 30, 30))], num_f=63, A=8, H=16, C=24, D=4, approx=True,
        note="35-qubit 5-to-1 distillation (Steane), 148 stabiliser terms, 15 det + 5 obs",
    ),
    "cfg3_surface_d5": dict(
        n_direct=121, comps=[], num_f=121, A=0, H=0, C=0, D=0, approx=False,
        note="rotated surface code d=5 r=5: 120 det + 1 obs, all direct (rank 1)",
    ),
    "cfg3p_rank1": dict(
        n_direct=0, comps=[(1, 6, (1, 1))] * 121, num_f=121, A=2, H=4, C=6, D=0, approx=False,
        note="rank-1 non-direct variant of cfg3: 121 one-output components",
    ),
    "cfg4_cultivation_d3": dict(
        n_direct=20, comps=[(10, 64, (93,) * 10 + (94,))], num_f=90, A=12, H=24, C=32, D=6, approx=False,
        note="d=3 cultivation, 1024 stabiliser terms, exact branch",
    ),
    "cfg5_distill85": dict(
        n_direct=40, comps=[(5, 120, (16, 20, 24, 28, 29, 30))], num_f=160, A=8, H=24, C=40, D=4, approx=True,
        note="85-qubit 5-to-1 distillation (ColorEncoder5), 147 stabiliser terms, 40 det + 5 obs",
    ),
}


def _masks(rng, G, T, P, density):
    m = (rng.random((G, T, P)) < density).astype(np.uint8)
    if P > 0 and T > 0:
        empty = ~m.any(axis=-1)
        gi, ti = np.nonzero(empty)
        m[gi, ti, rng.integers(0, P, size=len(gi))] = 1
    return m


def random_level(rng, G: int, P: int, A: int, H: int, C: int, D: int, *, approx: bool, density: float = 0.15,
                 power2_range=(-4, 0), shared_masks: bool = False) -> CompiledScalarGraphs:
    """One level with i.i.d. contents ("Synthetic inputs").  ``shared_masks``: every graph of the
    level uses the parameter masks of graph 0 (phases, constants and prefactors stay per graph) -- the shape a stabiliser
    decomposition of ONE ZX diagram produces, where the terms of a level differ in their phases, not in their wiring."""
    if shared_masks:
        draw = _masks

        def _shared(rng_, G_, T, P_, dens):
            return np.repeat(draw(rng_, 1, T, P_, dens), G_, axis=0)

        masks = _shared
    else:
        masks = _masks

    def counts(T):
        return rng.integers((T + 1) // 2, T + 1, size=G) if T else np.zeros(G, np.int64)

    aff = None
    if approx:
        theta = rng.uniform(0, 2 * np.pi, size=G)
        aff = (rng.uniform(0.5, 1.0, size=G) * np.exp(1j * theta)).astype(np.complex64)
    return make_scalar_graphs(
        P,
        num_graphs=G,
        node=(rng.integers(0, 8, (G, A)), masks(rng, G, A, P, density), counts(A)),
        halfpi=(rng.choice([2, 4, 6], size=(G, H)), masks(rng, G, H, P, density)),
        pi=(rng.integers(0, 2, (G, C)), masks(rng, G, C, P, density), rng.integers(0, 2, (G, C)), masks(rng, G, C, P, density)),
        pairs=(rng.integers(0, 8, (G, D)), masks(rng, G, D, P, density), rng.integers(0, 8, (G, D)), masks(rng, G, D, P, density), counts(D)),
        phase_indices=rng.integers(0, 8, G),
        floatfactor=np.tile(np.array([1, 0, 0, 0]), (G, 1)),
        power2=rng.integers(power2_range[0], power2_range[1] + 1, G),
        approximate_floatfactors=aff,
        has_approximate_floatfactors=approx,
    )


def _float_amplitude(lv: CompiledScalarGraphs, x: np.ndarray) -> complex:
    """Plain float64 value of one level at one parameter vector (closed forms of terms.py; used only
    to pick the per-level rescaling, never as a reference)."""
    if lv.num_graphs == 0:
        return 0.0
    x = x.astype(np.int64)
    w = np.exp(1j * np.pi / 4)
    n, h, p, q, pre = lv.node_phases, lv.halfpi_phases, lv.pi_products, lv.phase_pairs, lv.prefactor

    def par(m):
        return (m.astype(np.int64) @ x) % 2

    A, D = n.phases.shape[1], q.alpha.shape[1]
    node = np.where(np.arange(A)[None] < n.counts[:, None], 1 + w ** ((4 * par(n.params) + n.phases) % 8), 1).prod(axis=1)
    hp = w ** ((par(h.params) * h.coeffs).sum(axis=1) % 8)
    pi = (-1.0) ** (((p.psi_const + par(p.psi_params)) % 2 * ((p.phi_const + par(p.phi_params)) % 2)).sum(axis=1) % 2)
    ea = w ** ((q.alpha + 4 * par(q.alpha_params)) % 8)
    eb = w ** ((q.beta + 4 * par(q.beta_params)) % 8)
    pairs = np.where(np.arange(D)[None] < q.counts[:, None], 1 + ea + eb - ea * eb, 1).prod(axis=1)
    ff = pre.floatfactor.astype(np.float64) @ np.array([1, w, 1j, np.conj(w)])
    val = node * hp * pi * pairs * w ** pre.phase_indices.astype(np.int64) * ff * 2.0 ** pre.power2.astype(np.float64)
    if pre.has_approximate_floatfactors:
        val = val * pre.approximate_floatfactors.astype(np.complex128)
    return complex(val.sum())


def _rescale_levels(levels: list[CompiledScalarGraphs], F: int) -> None:
    """Shift each level's power2 so |E_k| / |E_{k-1}| is about 1/2 along the all-ones outcome path of the
    noiseless shot (in place)."""
    prev = None
    for k, lv in enumerate(levels):
        x = np.zeros(F + k, dtype=np.uint8)
        x[F:] = 1
        mag = abs(_float_amplitude(lv, x))
        if mag == 0 or not np.isfinite(mag):
            # a vanishing amplitude on this path: leave the level alone
            continue
        target = 1.0 if prev is None else prev * 0.5
        shift = int(np.round(np.log2(target / mag)))
        lv.prefactor.power2 += np.int32(shift)
        prev = mag * 2.0**shift


def synthetic_component(rng, n_c: int, F: int, graphs, f_pool, first_output: int, *, A, H, C, D, approx, density=0.15,
                        shared_masks: bool = False):
    levels = [random_level(rng, graphs[k], F + k, A, H, C, D, approx=approx, density=density, shared_masks=shared_masks)
              for k in range(n_c + 1)]
    _rescale_levels(levels, F)
    f_selection = np.sort(rng.choice(f_pool, size=F, replace=False)).astype(np.int32)
    return CompiledComponent(tuple(range(first_output, first_output + n_c)), f_selection, tuple(levels))


def synthetic_program(name: str, seed: int = 20260101, *, density: float = 0.15, shared_masks: bool = False,
                      graph_scale: int = 1) -> CompiledProgram:
    """Program with the shapes of benchmark configuration ``name`` (see ``CONFIGS``).  The keyword arguments vary the
    structure for sensitivity sweeps (tools/structure_sweep.py): mask density, masks shared by the graphs of a level,
    ``graph_scale`` times as many graphs per level; the defaults are the benchmark programs."""
    cfg = CONFIGS[name]
    rng = np.random.default_rng(seed)
    num_f = cfg["num_f"]
    n_direct = cfg["n_direct"]
    direct_f = rng.choice(num_f, size=n_direct, replace=False).astype(np.int32) if n_direct else np.zeros(0, np.int32)
    direct_flips = rng.integers(0, 2, n_direct).astype(bool) if n_direct else np.zeros(0, bool)
    comps = []
    out = n_direct
    for n_c, F, graphs in cfg["comps"]:
        comps.append(
            synthetic_component(
                rng, n_c, F, tuple(int(g) * graph_scale for g in graphs), np.arange(num_f), out, A=cfg["A"], H=cfg["H"], C=cfg["C"],
                D=cfg["D"], approx=cfg["approx"], density=density, shared_masks=shared_masks,
            )
        )
        out += n_c
    n_out = out
    # a non-trivial output permutation, as pipeline.py:90-99 produces (direct first, components by size)
    output_order = rng.permutation(n_out).astype(np.int32)
    prog = make_program(
        comps,
        direct_f_indices=direct_f,
        direct_flips=direct_flips,
        output_order=output_order,
        num_outputs=n_out,
        num_detectors=max(0, n_out - (5 if "distill" in name else 1)),
        num_f=num_f,
        meta={"config": name, "note": cfg["note"], "seed": seed, "synthetic": True},
    )
    return prog


def noise_probs(num_f: int, p: float = 1e-3) -> np.ndarray:
    """Per-f firing probability ``p * mult`` with ``mult`` cycling 1..15."""
    return p * (1 + (np.arange(num_f) % 15))
