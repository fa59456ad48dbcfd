"""Flatten a ``CompiledProgram`` into the bit-packed blob the CUDA library uploads to HBM.

The reference keeps every mask of g_{tki} as ``uint8[G, T, P]`` (one byte per
bit, ``compile/compile.py:40-334``) and contracts it with a float32 GEMM
(``utils/linalg.py:81-102``).  Here each mask row becomes ``W`` 32-bit words
(bit ``i`` of the parameter vector = bit ``i % 32`` of word ``i // 32``) and the
per-term constants ride in the same record, so one shared-memory read brings a
term's whole description.

Blob = ``uint32`` words: header (32 words) | direct table | component table |
level table | chunk table | f_selection | output destinations | chunk data.
A *chunk* is a 16-byte aligned run of whole graphs of one level, sized so a few
chunks fit the SM's shared memory; the kernel streams chunks with TMA bulk
copies (or keeps the whole data region resident when it fits).

Record layout per graph (``MODE_FAITHFUL``; ``W`` mask words each):

    node  slot  x A : mask[W], ctl = phase | valid << 3
    halfpi slot x H : mask[W], coeff
    pi    slot  x C : psi_mask[W], phi_mask[W], psi_const | phi_const << 1
    pair  slot  x D : alpha_mask[W], beta_mask[W], alpha | beta << 3 | valid << 6
    prefactor       : phase_idx, ff[4], power2, approx.re, approx.im (f32 bits)

See DESIGN.md for ``MODE_FAST``.
"""

from __future__ import annotations

import os

from dataclasses import dataclass, field

import numpy as np

from .program import CompiledProgram, CompiledScalarGraphs

MAGIC = 0x32425354  # "TSB2"
VERSION = 8
MODE_FAITHFUL = 0
MODE_FAST = 1
MODE_SLICED = 2
SLICED_AUTO_MAX_BYTES = 16 << 20  # mode="auto" keeps the per-row records beyond this size of the sliced data region

HEADER_WORDS = 32
COMP_WORDS = 8
LEVEL_WORDS = 12
CHUNK_WORDS = 4
PREFACTOR_WORDS = 8

# header slots
H_MAGIC, H_VERSION, H_MODE, H_W = 0, 1, 2, 3
H_NUM_F, H_N_OUT, H_N_DIRECT, H_N_COMP = 4, 5, 6, 7
H_N_DRAWS, H_N_LEVELS, H_N_CHUNKS, H_MAX_CHUNK = 8, 9, 10, 11
H_OFF_DIRECT, H_OFF_COMP, H_OFF_LEVEL, H_OFF_CHUNK = 12, 13, 14, 15
H_OFF_FSEL, H_OFF_DEST, H_OFF_DATA, H_DATA_WORDS = 16, 17, 18, 19
H_TOTAL_WORDS, H_WF64, H_WOUT64, H_OFF_TABLES = 20, 21, 22, 23
H_TABLE_WORDS = 24
H_ONE_ROW, H_ZERO_ROW = 25, 26
H_INDEX_SCALE = 28  # sliced: row index bytes are stored times this (see pack_sliced._index_words)
H_PLANE_ROWS = 27  # sliced: plane rows per graph slot in shared memory (vanish + index planes + multiplied-pair planes)


@dataclass
class PackedProgram:
    blob: np.ndarray  # uint32 [total_words]
    mode: int
    W: int
    num_f: int
    n_out: int
    n_components: int
    n_draws: int
    g_bytes: int  # algorithmic size of g_{tki} (SURVEY.md section 8(d) formula)
    stats: dict = field(default_factory=dict)

    @property
    def words_f64(self) -> int:
        return int(self.blob[H_WF64])

    @property
    def words_out64(self) -> int:
        return int(self.blob[H_WOUT64])


def pack_bits32(a: np.ndarray, W: int) -> np.ndarray:
    """``uint8[..., P]`` (0/1) -> ``uint32[..., W]`` little-endian bit order."""
    a = np.asarray(a, dtype=np.uint8) & 1
    P = a.shape[-1]
    pad = W * 32 - P
    if pad < 0:
        raise ValueError("mask wider than W words")
    if pad:
        a = np.concatenate([a, np.zeros(a.shape[:-1] + (pad,), np.uint8)], axis=-1)
    if a.size == 0:
        return np.zeros(a.shape[:-1] + (W,), dtype=np.uint32)
    by = np.packbits(a, axis=-1, bitorder="little")
    return np.ascontiguousarray(by).view(np.uint32).reshape(a.shape[:-1] + (W,))


def _f32_bits(x) -> np.ndarray:
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def algorithmic_g_bytes(program: CompiledProgram) -> int:
    """``sum G*[(A+H+2C+2D)*(8*W64+1) + 8 + 29]`` over all levels."""
    total = 0
    for comp in program.components:
        for lv in comp.compiled_scalar_graphs:
            G, P = lv.num_graphs, lv.n_params
            A = lv.node_phases.phases.shape[1]
            H = lv.halfpi_phases.coeffs.shape[1]
            C = lv.pi_products.psi_const.shape[1]
            D = lv.phase_pairs.alpha.shape[1]
            w64 = (P + 63) // 64
            total += G * ((A + H + 2 * C + 2 * D) * (8 * w64 + 1) + 8 + 29)
    return int(total)


# ------------------------------------------------------------------------------------------------
# exactness analysis: when may the device reorder the exact arithmetic?
# ------------------------------------------------------------------------------------------------


def _embedding_bound(ff: np.ndarray) -> np.ndarray:
    """Per row: max over the embeddings of Z[w] of |c0 + c1 w + c2 i + c3 conj(w)|.

    The Galois map w -> w^3 sends (1, w, i, conj(w)) to (1, w^3, -i, -w); the remaining two
    embeddings are complex conjugates of these.
    """
    ff = ff.astype(np.float64)
    r2 = np.sqrt(0.5)
    c0, c1, c2, c3 = ff[:, 0], ff[:, 1], ff[:, 2], ff[:, 3]
    m1 = np.hypot(c0 + r2 * (c1 + c3), c2 + r2 * (c1 - c3))
    m2 = np.hypot(c0 - r2 * (c1 + c3), r2 * (c1 - c3) - c2)
    return np.maximum(m1, m2)


def reorder_is_exact(program: CompiledProgram) -> tuple[bool, dict]:
    """Static sufficient condition for "any exact evaluation order gives the reference's bits".

    Every intermediate coefficient of the reference's int32 pipeline is bounded by the largest
    absolute value of the quantity over the embeddings of Z[w]; a node factor ``1 + w^k`` is at
    most 2 there, a pair factor at most ``2*sqrt(2)``.  If for every level
    ``sum_g 2^A_g * (2 sqrt2)^D_g * |ff_g| * 2^(power2_g - min power2) < 2^31`` then neither the
    products nor the power-aligned sums can wrap in *any* order, so every order yields the same
    canonical ``(coeffs, power)`` (``exact_scalar.py:124-137`` reduces to a fixpoint).
    """
    worst = 0.0
    for comp in program.components:
        for lv in comp.compiled_scalar_graphs:
            if lv.num_graphs == 0:
                continue
            A = np.minimum(lv.node_phases.counts, lv.node_phases.phases.shape[1]).astype(np.float64)
            D = np.minimum(lv.phase_pairs.counts, lv.phase_pairs.alpha.shape[1]).astype(np.float64)
            ffb = _embedding_bound(lv.prefactor.floatfactor)
            p2 = lv.prefactor.power2.astype(np.float64)
            log_term = A + 1.5 * D + np.log2(np.maximum(ffb, 1e-300))
            per_graph = float(np.max(log_term)) if len(log_term) else 0.0
            if lv.prefactor.has_approximate_floatfactors:
                level = per_graph  # no exact sum in the approximate branch
            else:
                level = float(np.log2(np.sum(np.exp2(log_term + p2 - p2.min()))))
            worst = max(worst, per_graph, level)
    return worst < 30.5, {"log2_bound": worst}


# ------------------------------------------------------------------------------------------------
# faithful records
# ------------------------------------------------------------------------------------------------


def _faithful_level_records(lv: CompiledScalarGraphs, W: int) -> np.ndarray:
    """-> uint32 [G, stride] records of one level."""
    G = lv.num_graphs
    n, h, p, q, pre = lv.node_phases, lv.halfpi_phases, lv.pi_products, lv.phase_pairs, lv.prefactor
    A, H, C, D = n.phases.shape[1], h.coeffs.shape[1], p.psi_const.shape[1], q.alpha.shape[1]
    parts = []
    # node
    valid = (np.arange(A)[None, :] < n.counts[:, None]).astype(np.uint32)
    ctl = (n.phases.astype(np.uint32) & 7) | (valid << 3)
    parts.append(np.concatenate([pack_bits32(n.params, W), ctl[..., None]], axis=-1).reshape(G, A * (W + 1)))
    # half-pi
    parts.append(
        np.concatenate([pack_bits32(h.params, W), h.coeffs.astype(np.uint32)[..., None]], axis=-1).reshape(G, H * (W + 1))
    )
    # pi products
    cst = (p.psi_const.astype(np.uint32) & 1) | ((p.phi_const.astype(np.uint32) & 1) << 1)
    parts.append(
        np.concatenate([pack_bits32(p.psi_params, W), pack_bits32(p.phi_params, W), cst[..., None]], axis=-1).reshape(
            G, C * (2 * W + 1)
        )
    )
    # phase pairs
    valid = (np.arange(D)[None, :] < q.counts[:, None]).astype(np.uint32)
    ctl = (q.alpha.astype(np.uint32) & 7) | ((q.beta.astype(np.uint32) & 7) << 3) | (valid << 6)
    parts.append(
        np.concatenate([pack_bits32(q.alpha_params, W), pack_bits32(q.beta_params, W), ctl[..., None]], axis=-1).reshape(
            G, D * (2 * W + 1)
        )
    )
    # prefactor
    pf = np.zeros((G, PREFACTOR_WORDS), dtype=np.uint32)
    pf[:, 0] = pre.phase_indices.astype(np.uint32) & 7
    pf[:, 1:5] = pre.floatfactor.astype(np.int32).view(np.uint32)
    pf[:, 5] = pre.power2.astype(np.int32).view(np.uint32)
    pf[:, 6] = _f32_bits(pre.approximate_floatfactors.real)
    pf[:, 7] = _f32_bits(pre.approximate_floatfactors.imag)
    parts.append(pf)
    return np.ascontiguousarray(np.concatenate(parts, axis=1), dtype=np.uint32)


# ------------------------------------------------------------------------------------------------
# blob assembly
# ------------------------------------------------------------------------------------------------


def pack_program(
    program: CompiledProgram,
    *,
    mode: str | int = "auto",
    max_chunk_words: int = 8192,
    joint: bool = False,
    sliced_budget_bytes: int | None = None,
) -> PackedProgram:
    """Build the device blob.  ``mode``: "faithful", "fast", "sliced", "rowwise" (fast when provably exact, else
    faithful) or "auto" (sliced when provably exact and compact, else as "rowwise").

    ``joint=True`` accepts the two-level programs of ``CompiledStateProbs`` (``mode="joint"`` in
    ``compile_program``: level 1 plugs all outputs at once); such a blob can only be *evaluated*.
    """
    exact_ok, bound_info = reorder_is_exact(program)
    if mode == "auto" and exact_ok and not joint:
        # fastest first: the bit-sliced records, unless their decode tables blow up (graphs with many general phase
        # pairs) -- every CTA streams the whole data region once per batch
        try:
            pp = pack_program(program, mode="sliced", max_chunk_words=max_chunk_words, sliced_budget_bytes=SLICED_AUTO_MAX_BYTES)
            if pp.stats["data_bytes"] <= SLICED_AUTO_MAX_BYTES:
                return pp
        except ValueError:
            pass
    if mode in ("auto", "rowwise"):
        if exact_ok:
            try:
                return pack_program(program, mode="fast", max_chunk_words=max_chunk_words, joint=joint)
            except ValueError:
                pass  # field ranges of the packed accumulator exceeded: keep the reference's order
        mode_id = MODE_FAITHFUL
    elif mode in ("faithful", MODE_FAITHFUL):
        mode_id = MODE_FAITHFUL
    elif mode in ("fast", MODE_FAST, "sliced", MODE_SLICED):
        if not exact_ok:
            raise ValueError(
                f"fast mode is not provably exact for this program (log2 bound {bound_info['log2_bound']:.1f} >= 30.5)"
            )
        mode_id = MODE_SLICED if mode in ("sliced", MODE_SLICED) else MODE_FAST
    else:
        raise ValueError(f"unknown mode {mode!r}")

    num_f = program.infer_num_f()
    n_out = int(program.num_outputs)
    n_direct = len(program.direct_f_indices)
    comps = program.components
    n_draws = sum(len(c.compiled_scalar_graphs) - 1 for c in comps)
    if not joint and n_direct + n_draws != n_out and n_out != 0:
        raise ValueError("num_outputs does not match direct + compiled outputs")

    # parameter words per shot: widest level over all components (+1 constant-one bit in fast mode)
    extra = 1 if mode_id == MODE_FAST else 0
    if joint and mode_id == MODE_SLICED:
        raise ValueError("joint-mode programs are evaluated per row: use the fast or faithful records")
    max_p = 0
    for c in comps:
        for lv in c.compiled_scalar_graphs:
            max_p = max(max_p, lv.n_params)
        F = len(c.f_selection)
        n_c = len(c.compiled_scalar_graphs) - 1
        for k, lv in enumerate(c.compiled_scalar_graphs):
            want = F + k
            if not joint and lv.n_params != want:
                raise ValueError(f"level {k} of a component has n_params={lv.n_params}, expected F_c + k = {want}")
        max_p = max(max_p, F + n_c)
    W = max(1, (max_p + extra + 31) // 32)

    # destination column of every combined column (direct first, then components in order)
    if program.output_reindex is not None:
        reindex = np.asarray(program.output_reindex, dtype=np.int64)
        dest = np.empty(n_out, dtype=np.int64)
        dest[reindex] = np.arange(n_out)
    else:
        dest = np.arange(n_out, dtype=np.int64)

    direct_tab = np.zeros((n_direct, 2), dtype=np.uint32)
    if n_direct:
        direct_tab[:, 0] = np.asarray(program.direct_f_indices, dtype=np.uint32)
        direct_tab[:, 1] = dest[:n_direct].astype(np.uint32) | (np.asarray(program.direct_flips, dtype=np.uint32) << 31)
        if direct_tab[:, 0].max(initial=0) >= max(num_f, 1):
            raise ValueError("direct_f_indices out of range")

    comp_tab = np.zeros((len(comps), COMP_WORDS), dtype=np.uint32)
    level_rows = []
    chunk_rows = []
    data_parts = []
    fsel_parts = []
    data_off = 0
    fsel_off = 0
    draw = 0
    plane_rows = 12
    # Stage size of the sliced kernel's TMA ring.  Programs with exact levels run 8 warps per slab group with wider plane
    # buffers (multiplied pairs): 44 KB stages leave shared memory for a third group per SM (16 -> 24 warps; the kernel is
    # latency bound) and still hold a whole wave of eight graphs.  Half-size stages were measured slower (cfg4: 12.2 ->
    # 14.7 ms): four graphs per chunk idle half of every wave.
    has_exact_level = any(
        lv.num_graphs > 0 and not lv.prefactor.has_approximate_floatfactors for c in comps for lv in c.compiled_scalar_graphs
    )
    # rows = max_p + 2 (parameters, all-ones row, all-zeros row); doubled index bytes must stay below 256
    index_scale = 2 if max_p + 1 <= 127 else 1
    sliced_chunk_words = 11264 if has_exact_level else 12288
    if not has_exact_level:
        # All-approximate programs run seven groups of four warps per SM (sliced_kernels.cuh): the stage must leave room for
        # their matrices and plane buffers next to a two-stage ring in the 227 KB of an SM, or the launch plan falls back to
        # fewer groups in several rounds (measured: 0.68 -> 0.85 ms on a variant of cfg2 whose largest chunk was 1 KB bigger).
        room = 232448 - 256 - 7 * ((max_p + 2) * 32 + 2 * 4 * 12 * 32) * 4
        if room // 8 >= 8192:
            sliced_chunk_words = min(sliced_chunk_words, (room // 8) & ~31)
    if os.environ.get("TSIM_B200_SLICED_CHUNK_WORDS"):  # tuning knob: stage size of the ring (multiple of 32 words)
        sliced_chunk_words = max(1024, int(os.environ["TSIM_B200_SLICED_CHUNK_WORDS"]) & ~31)
    sliced_compact = False
    if mode_id == MODE_SLICED:
        from .pack_sliced import compact_items_pay

        sliced_compact = compact_items_pay([lv for c in comps for lv in c.compiled_scalar_graphs])
        if os.environ.get("TSIM_B200_SLICED_COMPACT"):  # tuning knob
            sliced_compact = os.environ["TSIM_B200_SLICED_COMPACT"] != "0"
    for ci, c in enumerate(comps):
        F = len(c.f_selection)
        n_c = len(c.compiled_scalar_graphs) - 1
        if F and int(np.max(c.f_selection)) >= num_f:
            raise ValueError("f_selection out of range")
        comp_tab[ci] = [F, n_c, fsel_off, draw, len(level_rows), n_c + 1, 0, 0]
        fsel_parts.append(np.asarray(c.f_selection, dtype=np.uint32))
        fsel_off += F
        draw += n_c
        for k, lv in enumerate(c.compiled_scalar_graphs):
            if mode_id == MODE_FAITHFUL:
                recs = _faithful_level_records(lv, W)
                A, H = lv.node_phases.phases.shape[1], lv.halfpi_phases.coeffs.shape[1]
                C, D = lv.pi_products.psi_const.shape[1], lv.phase_pairs.alpha.shape[1]
                graph_lists = [recs[g] for g in range(lv.num_graphs)]
                p_lo = 0
            elif mode_id == MODE_SLICED:
                from .pack_sliced import sliced_level_chunks, sliced_level_records

                left = None if sliced_budget_bytes is None else max(0, sliced_budget_bytes // 4 - data_off)
                graphs, (A, H, C, D), p_lo = sliced_level_records(lv, max_p, max_p + 1, budget_words=left, index_scale=index_scale,
                                                                  compact=sliced_compact)
                graph_lists = []
                for rec, _tbl in graphs:
                    plane_rows = max(plane_rows, 1 + (int(rec[1]) & 0xFF) + 2 * ((int(rec[1]) >> 16) & 0xFF))
            else:
                from .pack_fast import fast_level_records  # local import: keeps this module lean

                graph_lists, (A, H, C, D), p_lo = fast_level_records(lv, W, lv.n_params)
            first_chunk = len(chunk_rows)
            # greedy split of whole graphs into chunks of at most max_chunk_words
            cur, cur_words, cur_graphs = [], 0, 0
            def flush():
                nonlocal cur, cur_words, cur_graphs, data_off
                if not cur_graphs:
                    return
                arr = np.concatenate(cur) if cur else np.zeros(0, np.uint32)
                pad = (-len(arr)) % 4
                if pad:
                    arr = np.concatenate([arr, np.zeros(pad, np.uint32)])
                chunk_rows.append([data_off, len(arr), cur_graphs, 0])
                data_parts.append(arr)
                data_off += len(arr)
                cur, cur_words, cur_graphs = [], 0, 0
            for rec in graph_lists:
                if cur_graphs and cur_words + len(rec) > max_chunk_words:
                    flush()
                cur.append(rec)
                cur_words += len(rec)
                cur_graphs += 1
            flush()
            if mode_id == MODE_SLICED:
                # chunk = directory | records | decode tables (pack_sliced.py); offsets inside are chunk-relative
                for arr, n_graphs in sliced_level_chunks(graphs, sliced_chunk_words):
                    chunk_rows.append([data_off, len(arr), n_graphs, 0])
                    data_parts.append(arr)
                    data_off += len(arr)
            flags = 1 if lv.prefactor.has_approximate_floatfactors else 0
            level_rows.append(
                [lv.num_graphs, lv.n_params, A, H, C, D, flags, first_chunk, len(chunk_rows) - first_chunk,
                 int(np.int32(p_lo).view(np.uint32)), 0, 0]
            )

    level_tab = np.asarray(level_rows, dtype=np.uint32).reshape(-1, LEVEL_WORDS)
    chunk_tab = np.asarray(chunk_rows, dtype=np.uint32).reshape(-1, CHUNK_WORDS)
    fsel = np.concatenate(fsel_parts) if fsel_parts else np.zeros(0, np.uint32)
    dest_tab = dest[n_direct:].astype(np.uint32) if not joint else np.zeros(n_draws, np.uint32)
    data = np.concatenate(data_parts) if data_parts else np.zeros(0, np.uint32)

    header = np.zeros(HEADER_WORDS, dtype=np.uint32)
    off = HEADER_WORDS
    header[H_OFF_TABLES] = off
    header[H_OFF_DIRECT] = off
    off += direct_tab.size
    header[H_OFF_COMP] = off
    off += comp_tab.size
    header[H_OFF_LEVEL] = off
    off += level_tab.size
    header[H_OFF_CHUNK] = off
    off += chunk_tab.size
    header[H_OFF_FSEL] = off
    off += fsel.size
    header[H_OFF_DEST] = off
    off += dest_tab.size
    header[H_TABLE_WORDS] = off - HEADER_WORDS
    off += (-off) % 32  # 128-byte align the data region
    header[H_OFF_DATA] = off
    header[H_DATA_WORDS] = data.size
    total = off + data.size
    header[H_TOTAL_WORDS] = total
    header[[H_MAGIC, H_VERSION, H_MODE, H_W]] = [MAGIC, VERSION, mode_id, W]
    header[[H_NUM_F, H_N_OUT, H_N_DIRECT, H_N_COMP]] = [num_f, n_out, n_direct, len(comps)]
    header[[H_N_DRAWS, H_N_LEVELS, H_N_CHUNKS]] = [n_draws, len(level_rows), len(chunk_rows)]
    header[H_MAX_CHUNK] = int(chunk_tab[:, 1].max()) if len(chunk_rows) else 0
    header[H_WF64] = max(1, (num_f + 63) // 64)
    header[H_WOUT64] = max(1, (n_out + 63) // 64)
    header[H_ONE_ROW] = max_p
    header[H_ZERO_ROW] = max_p + 1
    header[H_PLANE_ROWS] = plane_rows
    header[H_INDEX_SCALE] = index_scale if mode_id == MODE_SLICED else 1

    blob = np.zeros(total, dtype=np.uint32)
    blob[:HEADER_WORDS] = header
    pos = HEADER_WORDS
    for arr in (direct_tab, comp_tab, level_tab, chunk_tab, fsel, dest_tab):
        blob[pos : pos + arr.size] = arr.reshape(-1)
        pos += arr.size
    blob[int(header[H_OFF_DATA]) :] = data

    return PackedProgram(
        blob=blob,
        mode=mode_id,
        W=W,
        num_f=num_f,
        n_out=n_out,
        n_components=len(comps),
        n_draws=n_draws,
        g_bytes=algorithmic_g_bytes(program),
        stats={"joint": bool(joint), "reorder_exact": exact_ok, **bound_info, "data_bytes": int(data.size * 4), "n_chunks": len(chunk_rows)},
    )
