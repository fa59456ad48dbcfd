"""MODE_SLICED records: the same monoid arithmetic as MODE_FAST, evaluated bit-sliced.

In MODE_SLICED one thread owns a *slab* of 32 shots.  The slab's parameter matrix is kept transposed
(row i = bit i of the 32 shots), so a term's GF(2) contraction for all 32 shots is the XOR of the rows
selected by its mask -- no popcount, and a cost proportional to the mask's weight.  The exponents of the
monoid element ``w^a (1+sqrt2)^b`` (see ``pack_fast.py``) are accumulated as bit-planes (3 planes for
``a``, a bit-sliced counter for ``b``, one plane for "some factor vanished"); only the final decode of
a graph's value and the sum over graphs run per shot.

Graph record (uint32 words):

    [0]  n_terms | n_general_pairs << 16
    [1]  (b_base + 64) | n_b_planes << 8
    [2]  p_T   [3] power2   [4] shift (= p_T + power2 - p_lo)   [5] approx.re   [6] approx.im   [7] record words
    [8..11] K1 = w^a_static (1+w)^(n&3) floatfactor     [12..15] K2 = K1 sqrt2
    [16..17] ctl bytes of up to 8 general pairs (alpha | beta << 3)
    [18..19] reserved
    then the term stream, one control word per term followed by its index words:
        cw bits 0-1 type (0 LIN, 1 PI, 2 PAIR_GENERAL, 3 PAIR_MONOID), bits 2-7 n1, bits 8-13 n2 (index words of
        the first / second parity; an index word holds four row indices, padded with the all-zero row), bit 31 generic.
        Every term record is 16-byte aligned.  Compact form (all parities <= 3 index words, the common case):
            LIN  [cw, a0, a1, a2]                      others  [cw, a0, a1, a2, b0, b1, b2, extra]
        so a term is one or two 128-bit shared-memory reads at fixed positions (unused index words are never read).
        Generic form (bit 31): [cw, extra?, a..., b...] padded to a multiple of four words.
        LIN          bits 14-16 da, 17-18 bmode (1: count p, 2: count ~p), 19-20 zmode (1: Z |= p, 2: Z |= ~p)
        PI           A2 ^= p1 & p2
        PAIR_GENERAL bits 14-17 slot: parity words are kept for the per-shot ring product
        PAIR_MONOID  one extra word after cw: for v in (pa, pb, pa&pb): da_v (3 bits), db_v + 3 (3 bits); bits 18-21
                     truth table of vanishing combinations (bit pa + 2 pb)

Row indices: parameter i -> row i; row ``one_row`` is all ones (constants of the pi family), ``zero_row`` all zeros.
"""

from __future__ import annotations

import math

import numpy as np

from .pack_fast import MONOID, ONE_PLUS_W_POW, SQRT2, _zw_mul
from .program import CompiledScalarGraphs

SLICED_HEADER_WORDS = 20
MAX_GENERAL_PAIRS = 8
B_OFFSET = 64
T_LIN, T_PI, T_PAIR_GENERAL, T_PAIR_MONOID = 0, 1, 2, 3

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


GENERIC = 1 << 31


def _term_record(cw: int, i1: list[int], i2: list[int] | None, extra: int | None, zero_row: int) -> list[int]:
    """Compact (4 or 8 words) or generic (padded to 4) term record.  Unused index words of the compact form name
    the all-zero row four times: the kernel loads all twelve rows of a parity unconditionally."""
    two = i2 is not None
    zw = zero_row * 0x01010101
    if len(i1) <= 3 and (not two or len(i2) <= 3):
        rec = [cw] + i1 + [zw] * (3 - len(i1))
        if two:
            rec += i2 + [zw] * (3 - len(i2)) + [extra or 0]
        return rec
    rec = [cw | GENERIC] + ([extra] if extra is not None else []) + i1 + (i2 or [])
    return rec + [0] * ((-len(rec)) % 4)


def _index_words(rows: list[int], zero_row: int) -> list[int]:
    rows = list(rows)
    while len(rows) % 4:
        rows.append(zero_row)
    return [rows[i] | (rows[i + 1] << 8) | (rows[i + 2] << 16) | (rows[i + 3] << 24) for i in range(0, len(rows), 4)]


class _Unsupported(ValueError):
    pass


def sliced_level_records(lv: CompiledScalarGraphs, one_row: int, zero_row: int):
    """-> (list of per-graph uint32 records, (A, H, C, D), p_lo).  Raises ValueError if a graph does not fit."""
    G = lv.num_graphs
    n, h, p, q, pre = lv.node_phases, lv.halfpi_phases, lv.pi_products, lv.phase_pairs, lv.prefactor
    A, H, C, D = n.phases.shape[1], h.coeffs.shape[1], p.psi_const.shape[1], q.alpha.shape[1]
    approx = bool(pre.has_approximate_floatfactors)
    if zero_row > 255:
        raise _Unsupported("more than 254 parameters per level")

    def rows_of(mask, const=0):
        r = [int(i) for i in np.flatnonzero(mask)]
        if const & 1:
            r.append(one_row)
        return r

    recs, shifts_base = [], []
    for g in range(G):
        a_s = int(pre.phase_indices[g]) & 7
        b_base = 0
        n_tot = 0
        units = 0  # maximum value the b counter can reach
        terms: list[list[int]] = []
        always_zero = False

        def lin(rows, da=0, bmode=0, zmode=0):
            iw = _index_words(rows, zero_row)
            if len(iw) > 63:
                raise _Unsupported("mask too heavy")
            terms.append(_term_record(T_LIN | (len(iw) << 2) | ((da & 7) << 14) | (bmode << 17) | (zmode << 19), iw, None, None, zero_row))

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
            pc, fc = int(p.psi_const[g, j]) & 1, int(p.phi_const[g, j]) & 1
            r1, r2 = rows_of(p.psi_params[g, j], pc), rows_of(p.phi_params[g, j], fc)
            if not r1 or not r2:
                continue  # psi or phi is identically 0
            i1, i2 = _index_words(r1, zero_row), _index_words(r2, zero_row)
            if len(i1) > 63 or len(i2) > 63:
                raise _Unsupported("mask too heavy")
            terms.append(_term_record(T_PI | (len(i1) << 2) | (len(i2) << 8), i1, i2, None, zero_row))
        general_ctl = []
        for j in range(min(int(q.counts[g]), D)):
            al, be = int(q.alpha[g, j]) & 7, int(q.beta[g, j]) & 7
            r1, r2 = rows_of(q.alpha_params[g, j]), rows_of(q.beta_params[g, j])
            i1, i2 = _index_words(r1, zero_row), _index_words(r2, zero_row)
            if len(i1) > 63 or len(i2) > 63:
                raise _Unsupported("mask too heavy")
            combos = [monoid_exponents(pair_factor(al ^ (4 * pa), be ^ (4 * pb))) for pb in (0, 1) for pa in (0, 1)]
            nz = [c for c in combos if c not in (None, "zero")]
            monoid_ok = all(c is not None for c in combos) and nz and len({c[2] for c in nz}) == 1
            if not monoid_ok:
                if len(general_ctl) >= MAX_GENERAL_PAIRS:
                    raise _Unsupported("too many general phase pairs in one graph")
                slot = len(general_ctl)
                general_ctl.append(al | (be << 3))
                terms.append(_term_record(T_PAIR_GENERAL | (len(i1) << 2) | (len(i2) << 8) | (slot << 14), i1, i2, None, zero_row))
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
            terms.append(_term_record(T_PAIR_MONOID | (len(i1) << 2) | (len(i2) << 8), i1, i2, extra, zero_row))

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
        k1 = _zw_mul(_zw_mul(UNIT[a_s], ONE_PLUS_W_POW[r]), ff)
        k2 = _zw_mul(k1, SQRT2)
        if max(abs(v) for v in k1 + k2) >= 2**31:
            raise _Unsupported("graph constants overflow int32")
        if len(terms) > 0xFFFF:
            raise _Unsupported("too many terms")
        power2 = int(pre.power2[g])
        body = [w for t in terms for w in t] + [0] * 8  # slack: the kernel prefetches the next term's eight words
        words = np.zeros(SLICED_HEADER_WORDS + len(body), dtype=np.uint32)
        words[0] = len(terms) | (len(general_ctl) << 16)
        words[1] = (b_base + B_OFFSET) | (nb << 8)
        words[2] = np.int32(p_t).view(np.uint32)
        words[3] = np.int32(power2).view(np.uint32)
        aff = np.complex64(pre.approximate_floatfactors[g])
        words[5] = np.float32(aff.real).view(np.uint32)
        words[6] = np.float32(aff.imag).view(np.uint32)
        words[8:12] = np.array(k1, dtype=np.int64).astype(np.int32).view(np.uint32)
        words[12:16] = np.array(k2, dtype=np.int64).astype(np.int32).view(np.uint32)
        for s, ctl in enumerate(general_ctl):
            words[16 + s // 4] |= np.uint32(ctl << (8 * (s % 4)))
        words[SLICED_HEADER_WORDS:] = np.array(body, dtype=np.uint64).astype(np.uint32)
        pad = (-len(words)) % 4
        if pad:
            words = np.concatenate([words, np.zeros(pad, np.uint32)])
        words[7] = len(words)
        recs.append(words)
        shifts_base.append(p_t + power2)

    p_lo = min(shifts_base) if shifts_base else 0
    for words, sb in zip(recs, shifts_base):
        sh = sb - p_lo
        if not approx and sh > 30:
            raise _Unsupported("fixed-point shift exceeds 30 bits")
        words[4] = min(sh, 31)
    return recs, (A, H, C, D), p_lo
