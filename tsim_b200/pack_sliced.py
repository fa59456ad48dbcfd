"""MODE_SLICED records: the same monoid arithmetic as MODE_FAST, evaluated bit-sliced over shots.

A *slab* is 32 shots; its parameter matrix is kept transposed (row i = bit i of the 32 shots), so a term's GF(2)
contraction (``utils/linalg.py:81-102`` in the reference) for all 32 shots is the XOR of the rows selected by its mask --
no popcount, and a cost proportional to the mask's weight.  The exponents of the monoid element ``w^a (1+sqrt2)^b``
(see ``pack_fast.py``; term families ``compile/terms.py:56-187``) are accumulated as bit-planes (3 planes for ``a``, a
bit-sliced counter for ``b``, one plane for "some factor vanished").

The planes of a graph form an index ``a | cnt << 3 | pair parities << (3 + nb)`` into a per-graph *decode table*
built here at pack time: entry = the graph's contribution to the level sum for that index -- ``(re, im)`` float32 of
``to_complex(value) * approximate_floatfactor * 2^power2`` in the approximate branch (``compile/evaluate.py:56-59``; same
op order as the per-row kernels and the oracle, so the bits are identical), or the four int32 coefficients of
``value << shift`` in the exact branch (``:52-54``).  The device therefore never decodes a monoid element; per shot and
graph it gathers the index bits and adds one entry.

Pack-time algebra: every contribution that flips the top bit of ``a`` linearly in a parity (da & 4 of node / half-pi
terms, the cross terms ``pc phi' + fc psi'`` of the pi family's constants) is merged into one mask per graph (XOR of the
masks); half-pi terms are then ``a += 2 p``.  A graph with more general (odd-odd) phase pairs than the table index can
hold is split into variants, one per (pa, pb) combination of the excess pairs, gated so that exactly one contributes.

Chunk (the unit of the TMA stage ring) = directory (record offset of every graph, padded to 4 words) | graph
records | decode tables.  All offsets are in words relative to the chunk start.

Graph record (uint32 words):

    [0]  words of the term stream | n_general_pairs (in the table index) << 16
    [1]  n_index_bits | n_b_planes << 8
    [2]  offset of the decode table (filled in when the chunk is assembled)
    [3]  record words | words of the main part of the term stream << 16 (the rest, the aux part, holds pi runs only)
    [1]  also: number of *multiplied* general pairs << 16 (exact levels; their control bytes are the record's last 4 words)
    [4..7] zero (the decode entry of a shot whose value vanished)
    then the term stream as typed runs (``_emit_runs``): LIN / LIN2 / PI / PAIR items with straight-line parities of
    8 / 12 / 16 rows, and a generic block stream (``_block``) for heavier masks.  Ops:
        LIN      params da (3 bits) | bmode << 3 (1: count p, 2: count ~p) | zmode << 5 (1: Z |= p, 2: Z |= ~p)
        PI       A2 ^= q & p
        PAIRGEN  params slot: (q, p) become index planes of the decode table
        PAIRMON  params: for v in (q, p, q & p): da_v (3 bits), db_v + 3 (3 bits); bits 18-21 truth table of the
                 vanishing combinations (bit q + 2 p)
        FIRST    (generic stream only) keep the parity as q, the first half of a two-parity term

Row indices (8 bits, four per word): parameter i -> row i; ``zero_row`` is all zeros (padding); ``one_row`` is all ones
(kept for generic use: the pi constants are folded away at pack time).
"""

from __future__ import annotations

import math
import os

import numpy as np

from .pack_fast import MONOID, ONE_PLUS_W_POW, SQRT2, _zw_mul
from .program import CompiledScalarGraphs

SLICED_HEADER_WORDS = 8
MAX_GENERAL_PAIRS = 3
MAX_MUL_PAIRS = 4  # exact levels: general pairs applied as ring factors after the table lookup (two-stage decode)
MAX_INDEX_BITS = 11
MAX_CHUNK_WORDS = 12288  # 48 KB stages
B_OFFSET = 64

UNIT = [(1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, -1), (-1, 0, 0, 0), (0, -1, 0, 0), (0, 0, -1, 0), (0, 0, 0, 1)]
ONE_PLUS_SQRT2 = (1, 1, 0, 1)
SQRT2_MINUS_ONE = (-1, 1, 0, 1)
ONE_PLUS_W = (1, 1, 0, 0)


def _zw_pow(base, e):
    out = (1, 0, 0, 0)
    for _ in range(e):
        out = _zw_mul(out, base)
    return out


def _to_complex(c):
    w = complex(math.sqrt(0.5), math.sqrt(0.5))
    return c[0] + c[1] * w + c[2] * 1j + c[3] * w.conjugate()


def monoid_exponents(c):
    """``(a, b, n)`` with ``c == w^a (1+sqrt2)^b (1+w)^n`` exactly, or ``None`` (``c == 0`` -> ``"zero"``)."""
    c = tuple(int(v) for v in c)
    if not any(c):
        return "zero"
    # field norm = |v|^2 |sigma(v)|^2 must be a power of two
    a0, a1, a2, a3 = c
    # |v|^2 = X + Y sqrt2 with X = sum c_i^2, Y = c0 (c1 + c3) + c2 (c1 - c3)
    X = a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3
    Y = a0 * (a1 + a3) + a2 * (a1 - a3)
    norm = X * X - 2 * Y * Y
    if norm <= 0 or norm & (norm - 1):
        return None
    n = norm.bit_length() - 1
    v = _to_complex(c) / _to_complex(ONE_PLUS_W) ** n
    b = round(math.log(abs(v)) / math.log(1 + math.sqrt(2)))
    a = round(math.atan2(v.imag, v.real) / (math.pi / 4)) % 8
    cand = _zw_mul(_zw_mul(UNIT[a], _zw_pow(ONE_PLUS_SQRT2 if b >= 0 else SQRT2_MINUS_ONE, abs(b))), _zw_pow(ONE_PLUS_W, n))
    if cand != c:
        return None
    return a, b, n


def pair_factor(alpha: int, beta: int):
    ua, ub, uc = UNIT[alpha & 7], UNIT[beta & 7], UNIT[(alpha + beta) & 7]
    return tuple(int(i == 0) + ua[i] + ub[i] - uc[i] for i in range(4))


OP_FIRST, OP_LIN, OP_PI, OP_PAIRGEN, OP_PAIRMON = 0, 1, 2, 3, 4
CLASS_WORDS = (2, 3, 4)  # index words (four 8-bit row indices each) of the three straight-line parity classes
RUN_LIN = 0  # + class (0..2)
RUN_PI = 3  # + index of (class1, class2), class1 <= class2, in PI_CLASSES
RUN_LIN2 = 9  # + class: linear terms that only add 2 to ``a`` (the half-pi family after merging)
RUN_PAIR = 12  # + max(class1, class2): PAIRGEN / PAIRMON items
RUN_GENERIC = 15
PI_CLASSES = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
# one-word parities (<= 4 rows: sparse masks, the common case of structured programs) get compact items
RUN_LIN_1 = 16  # item = [params, index word]
RUN_LIN2_1 = 17  # item = [2, index word]
RUN_PI_1 = 18  # + (words of the heavier parity - 1), 0..2: item = [index word of the lighter parity, 3 index words of the heavier]
RUN_PAIR_1 = 21  # item = [op | params << 3, index word, index word, 0]
RUN_GENERIC_PI = 22  # a generic block stream that holds pi terms only (FIRST + PI block pairs): may live in the aux part


def _index_words(rows, n_words: int, zero_row: int, scale: int = 1) -> list[int]:
    """Row indices four per word (the kernel turns byte k into an address with one dp4a), padded with the zero row.
    Bytes hold ``row * scale`` (``H_INDEX_SCALE``: 2 when the program has at most 127 rows, so that the kernel's 64-bit
    lanes -- row stride 256 bytes -- are addressed with the same byte selectors)."""
    r = [v * scale for v in list(rows) + [zero_row] * (4 * n_words - len(rows))]
    return [r[i] | (r[i + 1] << 8) | (r[i + 2] << 16) | (r[i + 3] << 24) for i in range(0, 4 * n_words, 4)]


def _row_class(rows) -> int:
    for cls, nw in enumerate(CLASS_WORDS):
        if len(rows) <= 4 * nw:
            return cls
    return 3


def _block(op: int, params: int, rows: list[int], zero_row: int, scale: int = 1) -> list[int]:
    """Generic parity block: ``[op | params << 3, n_words, index words...]`` padded to an even number of words."""
    if params >> 29:
        raise _Unsupported("block parameters do not fit")
    n = max(1, (len(rows) + 3) // 4)
    n += n % 2
    return [op | (params << 3), n] + _index_words(rows, n, zero_row, scale)


def _emit_runs(terms, zero_row: int, scale: int = 1, rotate: int = 0, compact: bool = True) -> list[int]:
    """The term stream of a graph as typed runs: ``[kind | count << 16, 0, 0, 0]`` followed by ``count`` items of one
    shape, so that the kernel dispatches once per run and then loops over straight-line code.  The order of a graph's
    terms is irrelevant (their plane updates commute).  Runs and items are 16-byte aligned.

        LIN_1 / LIN2_1   item = [params, index word] (<= 4 rows); PI_1 + k: item = [index word, 3 index words] (lighter parity
                         <= 4 rows, heavier <= 4 (k + 1)); PAIR_1: item = [op | params << 3, index word, index word, 0]
        LIN + cls        item = [params, index words...]: 4 words for classes 0 / 1 (<= 8 / 12 rows), 8 for class 2 (<= 16)
        LIN2 + cls       the same for params == 2 (a += 2 p and nothing else): the kernel skips the generic update
        PAIR + cls       item = [op | params << 3, 0, 0, 0, 4 index words of the first parity, 4 of the second];
                         cls = the heavier class, used for both
        PI + pair index  item = [4 index words of the lighter parity, 4 of the heavier]; the class says how many are used
        GENERIC          count = words; items are ``_block`` streams (two-parity ops as a FIRST block + op block);
                         GENERIC_PI: the same for heavy pi terms alone
    """
    runs: dict[int, list[int]] = {}
    counts: dict[int, int] = {}
    heavy_pi: list[list[int]] = []
    zw = zero_row * scale * 0x01010101

    def add(kind, item):
        runs.setdefault(kind, []).extend(item)
        counts[kind] = counts.get(kind, 0) + 1

    def words_of(rows):
        return max(1, (len(rows) + 3) // 4)

    for t in terms:
        if t[0] == "lin":
            _, params, rows = t
            cls = _row_class(rows)
            if compact and words_of(rows) == 1:
                add(RUN_LIN2_1 if params == 2 else RUN_LIN_1, [params] + _index_words(rows, 1, zero_row, scale))
                continue
            if cls < 3:
                kind = (RUN_LIN2 if params == 2 else RUN_LIN) + cls
                iw = _index_words(rows, CLASS_WORDS[cls], zero_row, scale)
                add(kind, [params] + iw + [zw] * (3 - len(iw)) if cls < 2 else [params] + iw + [zw] * 3)
                continue
            words = _block(OP_LIN, params, rows, zero_row, scale)
        else:
            _, op, params, r1, r2 = t
            c1, c2 = _row_class(r1), _row_class(r2)
            if op == OP_PI and c1 < 3 and c2 < 3:
                if len(r1) > len(r2):
                    r1, r2, c1, c2 = r2, r1, c2, c1
                if compact and words_of(r1) == 1 and words_of(r2) <= 3:
                    add(RUN_PI_1 + words_of(r2) - 1, _index_words(r1, 1, zero_row, scale) + _index_words(r2, 3, zero_row, scale))
                    continue
                if c1 > c2:
                    r1, r2, c1, c2 = r2, r1, c2, c1
                add(RUN_PI + PI_CLASSES.index((c1, c2)), _index_words(r1, 4, zero_row, scale) + _index_words(r2, 4, zero_row, scale))
                continue
            if op in (OP_PAIRGEN, OP_PAIRMON) and c1 < 3 and c2 < 3:
                if compact and words_of(r1) == 1 and words_of(r2) == 1:
                    add(RUN_PAIR_1, [op | (params << 3)] + _index_words(r1, 1, zero_row, scale) + _index_words(r2, 1, zero_row, scale) + [0])
                    continue
                add(RUN_PAIR + max(c1, c2), [op | (params << 3), 0, 0, 0] + _index_words(r1, 4, zero_row, scale) + _index_words(r2, 4, zero_row, scale))
                continue
            words = _block(OP_FIRST, 0, r1, zero_row, scale) + _block(op, params, r2, zero_row, scale)
            if op == OP_PI:  # heavy pi terms get generic runs of their own: a helper warp can take them
                heavy_pi.append(words)
                continue
        runs.setdefault(RUN_GENERIC, []).extend(words)
        counts[RUN_GENERIC] = len(runs[RUN_GENERIC])
    for kind in (RUN_LIN_1, RUN_LIN2_1):  # two-word items: keep the next run header 16-byte aligned
        if kind in runs and len(runs[kind]) % 4:
            runs[kind] += [0, zw]  # a no-op item: params 0 (nothing to update) over the all-zeros row
            counts[kind] += 1
    body: list[int] = []
    # The order of the runs is free.  Graphs of a level often have the same mix of runs (always, when they share their
    # masks); starting every graph at a different run keeps the warps of a wave -- and the groups of a CTA, which walk
    # the same chunk -- out of step, so that their bursts of row loads do not all hit the LSU at the same time.
    order = sorted(runs)
    if order and os.environ.get("TSIM_B200_SLICED_ROTATE", "1") != "0":
        k = rotate % len(order)
        order = order[k:] + order[:k]
    for kind in order:
        if counts[kind] > 0xFFFF:
            raise _Unsupported("too many terms")
        if kind in (RUN_GENERIC, RUN_GENERIC_PI):
            runs[kind] += [0] * ((-len(runs[kind])) % 4)
            counts[kind] = len(runs[kind])
        body += [kind | (counts[kind] << 16), 0, 0, 0] + runs[kind]
    # heavy pi terms: runs of at most eight terms, so that the [main | aux] split can balance them
    for i in range(0, len(heavy_pi), 8):
        words = [w for t in heavy_pi[i : i + 8] for w in t]
        words += [0] * ((-len(words)) % 4)
        if len(words) > 0xFFFF:
            raise _Unsupported("too many terms")
        body += [RUN_GENERIC_PI | (len(words) << 16), 0, 0, 0] + words
    return body


def _run_loads(body: list[int]) -> list[tuple[int, int, int]]:
    """(kind, first word, row loads incl. padding) of every run of an emitted stream."""
    out, o = [], 0
    while o < len(body):
        kind, count = body[o] & 0xFFFF, body[o] >> 16
        start = o
        o += 4
        if kind < 3 or 9 <= kind < 12:
            nw = CLASS_WORDS[kind % 3]
            loads, o = 4 * nw * count, o + (4 if nw <= 3 else 8) * count
        elif 3 <= kind < 9:
            c1, c2 = PI_CLASSES[kind - 3]
            loads, o = 4 * (CLASS_WORDS[c1] + CLASS_WORDS[c2]) * count, o + 8 * count
        elif 12 <= kind < 15:
            loads, o = 8 * CLASS_WORDS[kind - 12] * count, o + 12 * count
        elif kind in (RUN_LIN_1, RUN_LIN2_1):
            loads, o = 4 * count, o + 2 * count
        elif RUN_PI_1 <= kind < RUN_PI_1 + 3:
            loads, o = 4 * (2 + kind - RUN_PI_1) * count, o + 4 * count
        elif kind == RUN_PAIR_1:
            loads, o = 8 * count, o + 4 * count
        else:
            loads, o = 4 * count, o + count  # generic block stream: count = words
        out.append((kind, start, loads))
    return out


AUX_SHARE = 0.6


def _is_pi_run(kind: int) -> bool:
    return RUN_PI <= kind < RUN_PI + 6 or RUN_PI_1 <= kind < RUN_PI_1 + 3 or kind == RUN_GENERIC_PI


def _split_streams(terms, zero_row: int, scale: int, *, rotate: int, compact: bool, split: bool):
    """-> (stream words, words of the main part).  The aux part (the tail) is a set of whole pi runs whose row loads come
    closest to ``AUX_SHARE`` of the graph's (the main warps also decode -- phase 2 -- while their helpers are already in
    the next wave, so the helpers take a little more than half); empty when ``split`` is off (exact levels: their
    kernels have no helper warps)."""
    body = _emit_runs(terms, zero_row, scale, rotate=rotate, compact=compact)
    if not split:
        return body, len(body)
    runs = _run_loads(body)
    total = sum(r[2] for r in runs)
    pi = sorted((r for r in runs if _is_pi_run(r[0])), key=lambda r: -r[2])
    aux_starts, acc = set(), 0
    for kind, start, loads in pi:  # largest first; take a run when it brings the aux part closer to one half
        if abs(acc + loads - total * AUX_SHARE) < abs(acc - total * AUX_SHARE):
            aux_starts.add(start)
            acc += loads
    bounds = [r[1] for r in runs] + [len(body)]
    main, aux = [], []
    for i, (kind, start, loads) in enumerate(runs):
        (aux if start in aux_starts else main).extend(body[start : bounds[i + 1]])
    return main + aux, len(main)


class _Unsupported(ValueError):
    pass


def compact_items_pay(levels) -> bool:
    """Whether the one-word item types (<= 4 rows per parity) are worth their extra runs: they are when at least a fifth
    of a program's parities have at most four rows.  Measured on variants of cfg2 (tools/structure_sweep.py): mask density
    0.05 (73 % of the parities) 0.58 -> 0.51 ms, density 0.15 (10 %) 0.66 -> 0.68 ms."""
    small = total = 0
    for lv in levels:
        for m in (lv.node_phases.params, lv.halfpi_phases.params, lv.pi_products.psi_params, lv.pi_products.phi_params,
                  lv.phase_pairs.alpha_params, lv.phase_pairs.beta_params):
            w = np.asarray(m).reshape(-1, m.shape[-1]).sum(axis=1) if m.size else np.zeros(0)
            total += int(np.count_nonzero(w))
            small += int(np.count_nonzero((w > 0) & (w <= 4)))
    return total > 0 and small * 5 >= total


def sliced_level_records(lv: CompiledScalarGraphs, one_row: int, zero_row: int, budget_words: int | None = None,
                         index_scale: int = 1, compact: bool = True):
    """-> (list of per-graph uint32 records, (A, H, C, D), p_lo).  Raises ValueError if a graph does not fit."""
    G = lv.num_graphs
    n, h, p, q, pre = lv.node_phases, lv.halfpi_phases, lv.pi_products, lv.phase_pairs, lv.prefactor
    A, H, C, D = n.phases.shape[1], h.coeffs.shape[1], p.psi_const.shape[1], q.alpha.shape[1]
    approx = bool(pre.has_approximate_floatfactors)
    if zero_row * index_scale > 255:
        raise _Unsupported("more than 254 parameters per level")

    def rows_of(mask, const=0):
        r = [int(i) for i in np.flatnonzero(mask)]
        if const & 1:
            r.append(one_row)
        return r

    recs, shifts_base, decode = [], [], []
    for g in range(G):
        a_s = int(pre.phase_indices[g]) & 7
        b_base = 0
        n_tot = 0
        units = 0  # maximum value the b counter can reach
        terms: list[list[int]] = []
        always_zero = False

        # Everything that flips the top bit of ``a`` linearly in a parity (da & 4 of a linear term, the cross terms of
        # the pi family's constants) is XOR-linear in the parameters: the masks are merged into one term per graph.
        m4: set[int] = set()

        def lin(rows, da=0, bmode=0, zmode=0):
            if da & 4:
                m4.symmetric_difference_update(rows)
                da &= 3
            if da or bmode or zmode:
                terms.append(("lin", (da & 7) | (bmode << 3) | (zmode << 5), rows))

        def two(op, params, r1, r2):
            terms.append(("two", op, params, r1, r2))

        for j in range(min(int(n.counts[g]), A)):
            ph = int(n.phases[g, j]) & 7
            rows = rows_of(n.params[g, j])
            base = MONOID[ph if ph != 4 else 0]
            flip = MONOID[(ph ^ 4) if (ph ^ 4) != 4 else 0]
            n_tot += base[2]
            if not rows:  # parity is always 0
                if ph == 4:
                    always_zero = True
                a_s += base[0]
                b_base += base[1]
                continue
            if ph in (0, 4):
                a_s += base[0]
                b_base += base[1]
                lin(rows, zmode=1 if ph == 0 else 2)
                continue
            da = (flip[0] - base[0]) & 7
            db = flip[1] - base[1]
            a_s += base[0]
            b_base += base[1]
            if db == 0:
                lin(rows, da=da)
            elif db == 1:
                lin(rows, da=da, bmode=1)
                units += 1
            elif db == -1:  # -p = (~p) - 1
                b_base -= 1
                lin(rows, da=da, bmode=2)
                units += 1
            else:
                raise _Unsupported("unexpected node exponent step")
        for j in range(H):
            c = int(h.coeffs[g, j]) & 7
            rows = rows_of(h.params[g, j])
            if c == 0 or not rows:
                continue
            lin(rows, da=c)
        for j in range(C):
            # (psi' + pc)(phi' + fc) = psi' phi' + pc phi' + fc psi' + pc fc: only the first product needs two parities
            pc, fc = int(p.psi_const[g, j]) & 1, int(p.phi_const[g, j]) & 1
            r1, r2 = rows_of(p.psi_params[g, j]), rows_of(p.phi_params[g, j])
            if pc:
                m4.symmetric_difference_update(r2)
            if fc:
                m4.symmetric_difference_update(r1)
            if pc and fc:
                a_s += 4
            if r1 and r2:
                two(OP_PI, 0, r1, r2)
        general = []
        for j in range(min(int(q.counts[g]), D)):
            al, be = int(q.alpha[g, j]) & 7, int(q.beta[g, j]) & 7
            r1, r2 = rows_of(q.alpha_params[g, j]), rows_of(q.beta_params[g, j])
            combos = [monoid_exponents(pair_factor(al ^ (4 * pa), be ^ (4 * pb))) for pb in (0, 1) for pa in (0, 1)]
            nz = [c for c in combos if c not in (None, "zero")]
            monoid_ok = all(c is not None for c in combos) and nz and len({c[2] for c in nz}) == 1
            if not monoid_ok:
                general.append((al, be, r1, r2))
                continue
            # value(pa, pb) as a polynomial: base + pa d10 + pb d01 + pa pb d11 in the exponents (a mod 8, b)
            ref = nz[0]
            e = [c if c != "zero" else ref for c in combos]  # exponents of vanishing combos are irrelevant
            ztt = sum(1 << i for i, c in enumerate(combos) if c == "zero")
            n_tot += ref[2]
            a00, b00 = e[0][0], e[0][1]
            d = [
                ((e[1][0] - a00) & 7, e[1][1] - b00),
                ((e[2][0] - a00) & 7, e[2][1] - b00),
                ((e[3][0] - e[1][0] - e[2][0] + a00) & 7, e[3][1] - e[1][1] - e[2][1] + b00),
            ]
            a_s += a00
            b_base += b00
            extra = 0
            for v, (da, db) in enumerate(d):
                if not -3 <= db <= 3:
                    raise _Unsupported("pair exponent step out of range")
                extra |= (da & 7) << (6 * v)
                extra |= (db + 3) << (6 * v + 3)
                if db < 0:
                    b_base += db  # each negative unit is counted as (~p) - 1
                units += abs(db)
            extra |= ztt << 18
            two(OP_PAIRMON, extra, r1, r2)

        if m4:
            # A2 ^= parity(m4) is linear, so a heavy merged mask goes out as pieces of at most 16 rows: they ride the
            # straight-line LIN runs instead of the generic block stream
            rows4 = sorted(m4)
            n_pieces = (len(rows4) + 15) // 16
            per = (len(rows4) + n_pieces - 1) // n_pieces
            for i in range(0, len(rows4), per):
                terms.append(("lin", 4, rows4[i : i + per]))
        if units > 31:
            raise _Unsupported("b counter needs more than 5 planes")
        nb = max(1, units.bit_length())
        p_t = n_tot >> 2
        r = n_tot & 3
        a_s = (a_s + 2 * p_t) & 7
        b_base += 2 * p_t
        if not (0 <= b_base + B_OFFSET and b_base + B_OFFSET + units <= 127):
            raise _Unsupported("Pell table range exceeded")
        ff = tuple(int(v) for v in pre.floatfactor[g])
        if always_zero:
            ff = (0, 0, 0, 0)
        power2 = int(pre.power2[g])
        # General pairs: the first few extend the decode-table index by (pa, pb); each further pair splits the graph
        # into one variant per (pa, pb) combination -- the pair's factor for that combination is folded into the
        # constants and a gate term makes the variant vanish for every other combination, so exactly one variant of
        # a graph contributes for a given shot (the float sum of the approximate branch sees the same addends).
        # Exact levels decode in two stages instead: the table holds the monoid part only, and the kernel multiplies the
        # looked-up value by each general pair's factor 1 + w^a + w^b - w^(a+b) (one shared 64-entry table, one ring
        # product per pair and shot; terms.py:164-187) -- wrapping Z[w] arithmetic is a ring, so the product equals the
        # single-table entry bit for bit, and the table no longer grows by 4x per pair.
        n_mul = 0 if approx else min(len(general), MAX_MUL_PAIRS)
        in_table = 0 if not approx else min(len(general), MAX_GENERAL_PAIRS, (MAX_INDEX_BITS - 3 - nb) // 2)
        excess = general[in_table + n_mul :]
        if len(excess) > 3:
            raise _Unsupported("too many general phase pairs in one graph")
        n_idx = 3 + nb + 2 * in_table
        general_ctl = []
        mul_ctl = []
        for slot, (al, be, r1, r2) in enumerate(general[: in_table + n_mul]):
            (general_ctl if approx else mul_ctl).append(al | (be << 3))
            two(OP_PAIRGEN, slot, r1, r2)  # planes 1 + 3 + nb + 2 slot (pa), + 1 (pb)
        for combo in range(4 ** len(excess)):
            ffv = ff
            gates = []
            for e, (al, be, r1, r2) in enumerate(excess):
                pa, pb = (combo >> (2 * e)) & 1, (combo >> (2 * e + 1)) & 1
                ffv = _zw_mul(ffv, pair_factor(al ^ (4 * pa), be ^ (4 * pb)))
                extra = sum(3 << (6 * v + 3) for v in range(3)) | ((0xF ^ (1 << (pa + 2 * pb))) << 18)
                gates.append(("two", OP_PAIRMON, extra, r1, r2))
            if excess and not any(ffv):
                continue  # this combination contributes nothing
            k1 = _zw_mul(_zw_mul(UNIT[a_s], ONE_PLUS_W_POW[r]), ffv)
            k2 = _zw_mul(k1, SQRT2)
            if max(abs(v) for v in k1 + k2) >= 2**31:
                raise _Unsupported("graph constants overflow int32")
            # The stream is laid out as [main | aux]: aux holds whole runs of pi terms (A2 ^= q & p is a pure XOR into
            # the top plane of ``a``, so it commutes with everything else) worth about half of the graph's row loads.
            # Thick launches walk the stream as one; thin launches (one group per SM, bound by the latency of one warp's
            # walk through a graph) give the aux part to a helper warp and XOR its plane in (sliced_kernels.cuh).
            body, main_words = _split_streams(terms + gates, zero_row, index_scale, rotate=len(recs), compact=compact,
                                              split=approx)
            if len(body) > 0xFFFF:
                raise _Unsupported("too many terms")
            trailer = [sum(c << (8 * i) for i, c in enumerate(mul_ctl)), 0, 0, 0] if mul_ctl else []
            words = np.zeros(SLICED_HEADER_WORDS + len(body), dtype=np.uint32)
            words[0] = len(body) | (len(general_ctl) << 16)
            words[1] = n_idx | (nb << 8) | (len(mul_ctl) << 16)
            words[SLICED_HEADER_WORDS:] = np.array(body, dtype=np.uint64).astype(np.uint32)
            pad = (-len(words)) % 4
            if pad:
                words = np.concatenate([words, np.zeros(pad, np.uint32)])
            if trailer:  # the last four words of the record: control bytes (alpha | beta << 3) of the multiplied pairs
                words = np.concatenate([words, np.array(trailer, dtype=np.uint32)])
            words[3] = len(words) | (main_words << 16)
            recs.append(words)
            shifts_base.append(p_t + power2)
            decode.append(
                dict(nb=nb, n_idx=n_idx, b64=b_base + B_OFFSET, k1=k1, k2=k2, ctl=list(general_ctl), p_t=p_t, power2=power2,
                     aff=np.complex64(pre.approximate_floatfactors[g]))
            )

    if not approx and len(recs) > 1:
        # Exact levels add integers, so the graph order is free (the approximate branch is a sequential float sum in graph
        # order, evaluate.py:56-59, and keeps it): sort by stream length so that the graphs of a wave -- one per warp, all
        # waiting at the wave's barrier for the longest -- cost about the same.
        order = sorted(range(len(recs)), key=lambda i: -len(recs[i]))
        recs = [recs[i] for i in order]
        shifts_base = [shifts_base[i] for i in order]
        decode = [decode[i] for i in order]
    p_lo = min(shifts_base) if shifts_base else 0
    if budget_words is not None:  # fail before the tables are built (mode="auto" gives up on bulky programs)
        need = sum(len(r) for r in recs) + sum((2 if approx else 4) << d["n_idx"] for d in decode)
        if need > budget_words:
            raise _Unsupported("sliced data region exceeds the size budget")
    tables = []
    for d, sb in zip(decode, shifts_base):
        sh = sb - p_lo
        if not approx and sh > 30:
            raise _Unsupported("fixed-point shift exceeds 30 bits")
        tables.append(decode_table(d, approx, min(sh, 31)))
    return list(zip(recs, tables)), (A, H, C, D), p_lo


# ------------------------------------------------------------------------------------------------
# decode tables
# ------------------------------------------------------------------------------------------------


def _pell_table() -> np.ndarray:
    """uint32 [128, 2]: (1+sqrt2)^(e-64) = P + Q sqrt2 with the device's wrapping recurrences."""
    t = np.zeros((128, 2), dtype=np.uint64)
    P, Q = 1, 0
    for e in range(64):
        t[64 + e] = (P, Q)
        P, Q = (P + 2 * Q) & 0xFFFFFFFF, (P + Q) & 0xFFFFFFFF
    P, Q = 1, 0
    for e in range(65):
        t[64 - e] = (P, Q)
        P, Q = (2 * Q - P) & 0xFFFFFFFF, (P - Q) & 0xFFFFFFFF
    return t.astype(np.uint32)


PELL = _pell_table()
PAIR_TABLE = np.array(
    [[v & 0xFFFFFFFF for v in pair_factor(i & 7, i >> 3)] for i in range(64)], dtype=np.uint64
).astype(np.uint32)  # index alpha | beta << 3
_SQRT1_2 = np.array([0x3F3504F3], dtype=np.uint32).view(np.float32)[0]


def _mul_u32(x, y):
    """Wrapping Z[w] product of uint32 [..., 4] arrays (exact_scalar.py:31-39)."""
    a1, b1, c1, d1 = (x[..., i] for i in range(4))
    a2, b2, c2, d2 = (y[..., i] for i in range(4))
    with np.errstate(over="ignore"):
        return np.stack(
            [
                a1 * a2 + b1 * d2 - c1 * c2 + d1 * b2,
                a1 * b2 + b1 * a2 + c1 * d2 + d1 * c2,
                a1 * c2 + b1 * b2 + c1 * a2 - d1 * d2,
                a1 * d2 - b1 * c2 - c1 * b2 + d1 * a2,
            ],
            axis=-1,
        ).astype(np.uint32)


def _pow2_f32(p: int) -> np.float32:
    from .pack_fast import _pow2_bits

    return np.array([_pow2_bits(int(p))], dtype=np.uint32).view(np.float32)[0]


def decode_values(d: dict) -> np.ndarray:
    """uint32 [2^n_idx, 4]: the graph's value (wrapping int32 coefficients, power p_T) for every plane index."""
    n_idx, nb = d["n_idx"], d["nb"]
    idx = np.arange(1 << n_idx, dtype=np.int64)
    a = idx & 7
    cnt = (idx >> 3) & ((1 << nb) - 1)
    pq = PELL[(d["b64"] + cnt) & 127]
    k1 = np.array([v & 0xFFFFFFFF for v in d["k1"]], dtype=np.uint64).astype(np.uint32)
    k2 = np.array([v & 0xFFFFFFFF for v in d["k2"]], dtype=np.uint64).astype(np.uint32)
    with np.errstate(over="ignore"):
        v = k1[None, :] * pq[:, 0:1] + k2[None, :] * pq[:, 1:2]
        neg = lambda t: (np.uint32(0) - t).astype(np.uint32)
        # multiply by w^a: w (c0,c1,c2,c3) = (c3, c0, c1, -c2); i (..) = (-c2, c3, c0, -c1)
        r1 = np.stack([v[:, 3], v[:, 0], v[:, 1], neg(v[:, 2])], axis=1)
        v = np.where((a & 1)[:, None] != 0, r1, v)
        r2 = np.stack([neg(v[:, 2]), v[:, 3], v[:, 0], neg(v[:, 1])], axis=1)
        v = np.where((a & 2)[:, None] != 0, r2, v)
        v = np.where((a & 4)[:, None] != 0, neg(v), v).astype(np.uint32)
    for slot, ctl in enumerate(d["ctl"]):
        pa = (idx >> (3 + nb + 2 * slot)) & 1
        pb = (idx >> (3 + nb + 2 * slot + 1)) & 1
        v = _mul_u32(v, PAIR_TABLE[(ctl ^ (pa << 2) ^ (pb << 5)) & 63])
    return v


def decode_table(d: dict, approx: bool, shift: int) -> np.ndarray:
    """uint32 [2^n_idx * (2 | 4)]: per-index contribution of the graph to the level sum (see the module docstring)."""
    v = decode_values(d)
    if not approx:
        with np.errstate(over="ignore"):
            return (v * np.uint32(1 << (shift & 31))).astype(np.uint32).reshape(-1)
    f = v.view(np.int32).astype(np.float32)
    s2 = _SQRT1_2
    sc, pw = _pow2_f32(d["p_t"]), _pow2_f32(d["power2"])
    are, aim = np.float32(d["aff"].real), np.float32(d["aff"].imag)
    with np.errstate(all="ignore"):
        t1, t3 = f[:, 1] * s2, f[:, 3] * s2
        tre = ((f[:, 0] + t1) + t3) * sc
        tim = ((t1 + f[:, 2]) - t3) * sc
        ure = tre * are - tim * aim
        uim = tre * aim + tim * are
        out = np.stack([ure * pw, uim * pw], axis=1).astype(np.float32)
    return np.ascontiguousarray(out).view(np.uint32).reshape(-1)


def sliced_level_chunks(graphs, max_chunk_words: int = MAX_CHUNK_WORDS):
    """Group (record, table) pairs of one level into chunks -> list of (uint32 chunk, n_graphs).

    Graph counts per chunk are kept at multiples of 8 where the stage size allows (the kernel hands the graphs of a
    chunk to its warps in waves of 4 or 8)."""

    def size(gs):
        return ((len(gs) + 3) & ~3) + sum(len(r) + ((len(t) + 3) & ~3) for r, t in gs)

    chunks, i = [], 0
    while i < len(graphs):
        take = 0
        while i + take < len(graphs) and size(graphs[i : i + take + 1]) <= max_chunk_words:
            take += 1
        if take == 0:
            raise _Unsupported("a graph does not fit a stage")
        if i + take < len(graphs):  # not the end of the level: keep whole waves together
            for unit in (8, 4, 2):
                if take >= unit:
                    take -= take % unit
                    break
        gs = graphs[i : i + take]
        n = len(gs)
        total = size(gs)
        chunk = np.zeros(total, dtype=np.uint32)
        pos = (n + 3) & ~3
        for j, (r, _t) in enumerate(gs):
            chunk[j] = pos
            chunk[pos : pos + len(r)] = r
            pos += len(r)
        for j, (_r, t) in enumerate(gs):
            chunk[int(chunk[j]) + 2] = pos
            chunk[pos : pos + len(t)] = t
            pos += (len(t) + 3) & ~3
        assert pos == total
        chunks.append((chunk, n))
        i += take
    return chunks
