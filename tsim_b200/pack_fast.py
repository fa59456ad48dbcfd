"""MODE_FAST records: the node / half-pi / pi families as one packed-integer accumulation.

Every node factor ``1 + w^k`` (``terms.py:66-73``), every half-pi phase ``w^c`` (``:104-107``), every
pi sign ``(-1)^(psi*phi)`` (``:136-144``) and the static phase ``w^phase`` (``evaluate.py:37``) lies in
the multiplicative monoid ``{ w^a (1+sqrt2)^b (1+w)^n } U {0}`` of Z[w]:

    k      0            1        2            3             4   5             6             7
    1+w^k  2            1+w      1+i          1+w^3         0   1-w           1-i           1+conj(w)
    (a,b,n) (6,-2,4)    (0,0,1)  (0,-1,2)     (1,-1,1)      -   (6,-1,1)      (6,-1,2)      (7,0,1)

so a graph's product of those families is described by three small integers that *add*:
``acc = acc0 + sum_j parity_j * delta_j`` with fields ``a`` (bits 29-31, natural wrap mod 8), ``b + 64``
(bits 16-28) and the number of vanishing factors (bits 0-15).  ``n`` does not depend on the parities
(flipping a parity maps k -> k ^ 4, which keeps n unless the factor vanishes), so the power of two
``n >> 2`` and the residue ``(1+w)^(n & 3)`` are folded into the graph's constants at pack time
(``(1+w)^4 = 2 w^2 (1+sqrt2)^2``).  The device turns ``(a, b)`` back into four coefficients with a
Pell-number table and a signed permutation.  A phase pair (``terms.py:174-187``) whose four possible factors
all lie in the monoid (at least one of alpha, beta even) becomes three virtual terms (pa, pb, pa and pb); the
odd-odd pairs contain other primes and stay a plain ring product.  The reordering is only used when ``pack.reorder_is_exact`` holds.
"""

from __future__ import annotations

import numpy as np

from .program import CompiledScalarGraphs


def monoid_exponents(c):
    from .pack_sliced import monoid_exponents as f

    return f(c)


def pair_factor(alpha, beta):
    from .pack_sliced import pair_factor as f

    return f(alpha, beta)

FAST_HEADER_WORDS = 16
B_OFFSET = 64

# (a, b, n) with 1 + w^k = w^a (1+sqrt2)^b (1+w)^n ; k = 4 vanishes
MONOID = {0: (6, -2, 4), 1: (0, 0, 1), 2: (0, -1, 2), 3: (1, -1, 1), 5: (6, -1, 1), 6: (6, -1, 2), 7: (7, 0, 1)}


def _pow2_bits(p: int) -> np.uint32:
    """float32 bit pattern of 2**p: denormals kept, overflow -> +inf, underflow -> 0 (same as the device's pow2_f32)."""
    if p > 127:
        return np.uint32(0x7F800000)
    if p >= -126:
        return np.uint32((p + 127) << 23)
    if p >= -149:
        return np.uint32(1 << (p + 149))
    return np.uint32(0)


def round4(n: int) -> int:
    return (n + 3) & ~3


def lin_stride(W: int) -> int:
    return 2 if W == 1 else round4(W + 1)


def pi_stride(W: int) -> int:
    return 2 if W == 1 else round4(2 * W)


def pair_stride(W: int) -> int:
    return round4(2 * W + 1)


def mpair_stride(W: int) -> int:
    return round4(2 * W + 3)


def _zw_mul(x, y):
    a1, b1, c1, d1 = x
    a2, b2, c2, d2 = y
    return (
        a1 * a2 + b1 * d2 - c1 * c2 + d1 * b2,
        a1 * b2 + b1 * a2 + c1 * d2 + d1 * c2,
        a1 * c2 + b1 * b2 + c1 * a2 - d1 * d2,
        a1 * d2 - b1 * c2 - c1 * b2 + d1 * a2,
    )


ONE_PLUS_W_POW = [(1, 0, 0, 0), (1, 1, 0, 0), (1, 2, 1, 0), (1, 3, 3, -1)]
SQRT2 = (0, 1, 0, 1)


def _pack_acc(a: int, b: int, z: int) -> int:
    return (((a & 7) << 29) + ((b & 0x1FFF) << 16) + z) & 0xFFFFFFFF


def _mask_words(bits: np.ndarray, W: int, const: int = 0) -> np.ndarray:
    """0/1 vector -> W uint32 words; ``const`` rides on the always-one parameter (bit 31 of word W-1)."""
    from .pack import pack_bits32

    w = pack_bits32(bits[None, :], W)[0].copy()
    if const & 1:
        w[W - 1] |= np.uint32(0x80000000)
    return w


def fast_level_records(lv: CompiledScalarGraphs, W: int, n_params: int):
    """-> (list of per-graph uint32 records, (A, H, C, D) kept for the level table) and sets ``p_lo``."""
    G = lv.num_graphs
    n, h, p, q, pre = lv.node_phases, lv.halfpi_phases, lv.pi_products, lv.phase_pairs, lv.prefactor
    A, H, C, D = n.phases.shape[1], h.coeffs.shape[1], p.psi_const.shape[1], q.alpha.shape[1]
    if n_params > 32 * W - 1:
        raise ValueError("no room for the always-one parameter bit")
    SL, SP, SD = lin_stride(W), pi_stride(W), pair_stride(W)
    approx = bool(pre.has_approximate_floatfactors)

    recs = []
    shifts_base = []
    for g in range(G):
        a0 = int(pre.phase_indices[g]) & 7
        b0 = 0
        z0 = 0
        n_tot = 0
        lin = []  # (mask words, delta)
        b_lo = b_hi = 0  # reachable range of the dynamic part of b
        z_hi = 0
        for j in range(min(int(n.counts[g]), A)):
            ph = int(n.phases[g, j]) & 7
            mask = n.params[g, j]
            s_par0, s_par1 = ph, ph ^ 4  # k for parity 0 / 1
            base = MONOID[s_par0 if s_par0 != 4 else 0]
            flip = MONOID[s_par1 if s_par1 != 4 else 0]
            a0 += base[0]
            b0 += base[1]
            n_tot += base[2]
            z_base = 1 if s_par0 == 4 else 0
            z_flip = 1 if s_par1 == 4 else 0
            z0 += z_base
            if not np.any(mask):
                continue  # parity is always 0: static
            da, db, dz = (flip[0] - base[0]) & 7, flip[1] - base[1], z_flip - z_base
            delta = (((da & 7) << 29) + (db << 16) + dz) & 0xFFFFFFFF
            lin.append((_mask_words(mask, W), delta))
            b_lo += min(db, 0)
            b_hi += max(db, 0)
            z_hi += max(dz, 0)
        for j in range(H):
            c = int(h.coeffs[g, j]) & 7
            mask = h.params[g, j]
            if c == 0 or not np.any(mask):
                continue
            lin.append((_mask_words(mask, W), (c << 29) & 0xFFFFFFFF))
        pis = []
        for j in range(C):
            pc, fc = int(p.psi_const[g, j]) & 1, int(p.phi_const[g, j]) & 1
            pm, fm = p.psi_params[g, j], p.phi_params[g, j]
            psi_static = not np.any(pm)
            phi_static = not np.any(fm)
            if (psi_static and pc == 0) or (phi_static and fc == 0):
                continue  # psi * phi == 0 always
            if psi_static and phi_static:
                a0 += 4  # both constants are 1
                continue
            pis.append((_mask_words(pm, W, pc), _mask_words(fm, W, fc)))
        pairs = []
        mpairs = []  # phase pairs whose four possible factors all lie in the monoid: three virtual linear terms
        for j in range(min(int(q.counts[g]), D)):
            al, be = int(q.alpha[g, j]) & 7, int(q.beta[g, j]) & 7
            ma, mb = _mask_words(q.alpha_params[g, j], W), _mask_words(q.beta_params[g, j], W)
            combos = [monoid_exponents(pair_factor(al ^ (4 * pa), be ^ (4 * pb))) for pb in (0, 1) for pa in (0, 1)]
            nz = [c for c in combos if c not in (None, "zero")]
            if not (all(c is not None for c in combos) and nz and len({c[2] for c in nz}) == 1):
                pairs.append((ma, mb, al | (be << 3)))
                continue
            e = [c if c != "zero" else nz[0] for c in combos]  # exponents of a vanishing factor are irrelevant
            zc = [1 if c == "zero" else 0 for c in combos]
            a0 += e[0][0]
            b0 += e[0][1]
            z0 += zc[0]
            n_tot += nz[0][2]
            d = []
            for hi, lo in (((1,), (0,)), ((2,), (0,)), ((3, 0), (1, 2))):
                da = (sum(e[i][0] for i in hi) - sum(e[i][0] for i in lo)) & 7
                db = sum(e[i][1] for i in hi) - sum(e[i][1] for i in lo)
                dz = sum(zc[i] for i in hi) - sum(zc[i] for i in lo)
                d.append((((da & 7) << 29) + (db << 16) + dz) & 0xFFFFFFFF)
            bs = [c[1] - e[0][1] for c in e]
            b_lo += min(bs)
            b_hi += max(bs)
            z_hi += max(z - zc[0] for z in zc)
            mpairs.append((ma, mb, d))

        p_t = n_tot >> 2
        r = n_tot & 3
        a0 += 2 * p_t
        b0 += 2 * p_t
        if not (0 <= b0 + B_OFFSET + b_lo and b0 + B_OFFSET + b_hi <= 127 and z0 + z_hi <= 0xFFFF):
            raise ValueError("graph exceeds the packed accumulator's field ranges")
        if len(lin) > 0xFFF or len(pis) > 0xFFF or len(pairs) > 0xFF or len(mpairs) > 0xFFFF:
            raise ValueError("too many terms in one graph for the fast record header")
        k1 = _zw_mul(ONE_PLUS_W_POW[r], tuple(int(v) for v in pre.floatfactor[g]))
        k2 = _zw_mul(k1, SQRT2)
        if max(abs(v) for v in k1 + k2) >= 2**31:
            raise ValueError("graph constants overflow int32")
        power2 = int(pre.power2[g])

        SM = mpair_stride(W)
        words = np.zeros(
            FAST_HEADER_WORDS + round4(len(lin) * SL) + round4(len(pis) * SP) + len(mpairs) * SM + len(pairs) * SD, dtype=np.uint32
        )
        words[7] = len(mpairs)
        words[0] = len(lin) | (len(pis) << 12) | (len(pairs) << 24)
        words[1] = _pack_acc(a0, b0 + B_OFFSET, z0)
        # exact float32 powers of two, precomputed (the approximate branch scales by 2^p_T and 2^power2 per graph)
        words[2] = _pow2_bits(p_t)
        words[3] = _pow2_bits(power2)
        aff = np.complex64(pre.approximate_floatfactors[g])
        words[5] = np.float32(aff.real).view(np.uint32)
        words[6] = np.float32(aff.imag).view(np.uint32)
        words[8:12] = np.array(k1, dtype=np.int64).astype(np.int32).view(np.uint32)
        words[12:16] = np.array(k2, dtype=np.int64).astype(np.int32).view(np.uint32)
        o = FAST_HEADER_WORDS
        for mask, delta in lin:
            words[o : o + W] = mask
            words[o + W] = delta
            o += SL
        o = FAST_HEADER_WORDS + round4(len(lin) * SL)
        for m1, m2 in pis:
            words[o : o + W] = m1
            words[o + W : o + 2 * W] = m2
            o += SP
        o = FAST_HEADER_WORDS + round4(len(lin) * SL) + round4(len(pis) * SP)
        for m1, m2, d in mpairs:
            words[o : o + W] = m1
            words[o + W : o + 2 * W] = m2
            words[o + 2 * W : o + 2 * W + 3] = d
            o += SM
        for m1, m2, ctl in pairs:
            words[o : o + W] = m1
            words[o + W : o + 2 * W] = m2
            words[o + 2 * W] = ctl
            o += SD
        recs.append(words)
        shifts_base.append(p_t + power2)

    p_lo = min(shifts_base) if shifts_base else 0
    for words, sb in zip(recs, shifts_base):
        sh = sb - p_lo
        if not approx and sh > 30:
            raise ValueError("fixed-point shift exceeds 30 bits")
        words[4] = min(sh, 31)
    return recs, (A, H, C, D), p_lo
