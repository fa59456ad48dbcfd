"""Pure-Python model of the MODE_FAST device decode (tsim_b200/csrc/sampler_kernels.cuh::eval_chunk_fast).

Reads the packed blob exactly as the kernel does (same record strides, same accumulator fields, same
Pell table, same signed-permutation rotation) with Python integers wrapped to 32 bits.  It lets the CPU
test-suite check the packer and the monoid arithmetic against the oracle without a GPU; it is not part
of the product.
"""

import numpy as np

from oracle.exact_scalar import pow2_f32, to_complex_parts
from tsim_b200 import pack as PK
from tsim_b200.pack_fast import FAST_HEADER_WORDS, lin_stride, mpair_stride, pair_stride, pi_stride, round4

M32 = 0xFFFFFFFF
W8 = np.exp(1j * np.pi / 4)


def _s32(v):
    v &= M32
    return v - (1 << 32) if v & 0x80000000 else v


def _mul(x, y):
    a1, b1, c1, d1 = x
    a2, b2, c2, d2 = y
    return tuple(
        v & M32
        for v in (
            a1 * a2 + b1 * d2 - c1 * c2 + d1 * b2,
            a1 * b2 + b1 * a2 + c1 * d2 + d1 * c2,
            a1 * c2 + b1 * b2 + c1 * a2 - d1 * d2,
            a1 * d2 - b1 * c2 - c1 * b2 + d1 * a2,
        )
    )


def _unit(k):
    k &= 7
    v = [0, 0, 0, 0]
    if k in (0, 4):
        v[0] = 1 if k == 0 else -1
    elif k in (1, 5):
        v[1] = 1 if k == 1 else -1
    elif k in (2, 6):
        v[2] = 1 if k == 2 else -1
    else:
        v[3] = -1 if k == 3 else 1
    return v


PAIR = []
for i in range(64):
    a, b = i & 7, i >> 3
    ua, ub, uc = _unit(a), _unit(b), _unit(a + b)
    PAIR.append(tuple((int(i2 == 0) + ua[i2] + ub[i2] - uc[i2]) & M32 for i2 in range(4)))

PELL = [None] * 128
P, Q = 1, 0
for e in range(64):
    PELL[64 + e] = (P & M32, Q & M32)
    P, Q = P + 2 * Q, P + Q
P, Q = 1, 0
for e in range(65):
    PELL[64 - e] = (P & M32, Q & M32)
    P, Q = 2 * Q - P, P - Q


def _rotate(v, a):
    c0, c1, c2, c3 = v
    if a & 1:
        c0, c1, c2, c3 = c3, c0, c1, -c2
    if a & 2:
        c0, c1, c2, c3 = -c2, c3, c0, -c1
    if a & 4:
        c0, c1, c2, c3 = -c0, -c1, -c2, -c3
    return tuple(v & M32 for v in (c0, c1, c2, c3))


def _par(xw, words):
    t = 0
    for a, b in zip(xw, words):
        t ^= int(a) & int(b)
    return bin(t).count("1") & 1


def evaluate_level(pp: PK.PackedProgram, comp: int, level: int, x_bits: np.ndarray):
    """-> ("exact", coeffs int32[4], power) or ("approx", re f32, im f32) for one parameter vector."""
    blob = pp.blob
    assert pp.mode == PK.MODE_FAST
    W = pp.W
    SL, SP, SD = lin_stride(W), pi_stride(W), pair_stride(W)
    comp_row = blob[int(blob[PK.H_OFF_COMP]) + comp * PK.COMP_WORDS :]
    lrow = int(comp_row[4]) + level
    lvl = blob[int(blob[PK.H_OFF_LEVEL]) + lrow * PK.LEVEL_WORDS :][: PK.LEVEL_WORDS]
    approx = bool(lvl[6] & 1)
    p_lo = int(np.uint32(lvl[9]).view(np.int32)) if hasattr(np.uint32(lvl[9]), "view") else int(lvl[9])
    p_lo = _s32(int(lvl[9]))
    xw = PK.pack_bits32(np.asarray(x_bits, np.uint8)[None, :], W)[0].astype(np.uint64)
    xw[W - 1] |= 0x80000000
    data = blob[int(blob[PK.H_OFF_DATA]) :]
    chunks = blob[int(blob[PK.H_OFF_CHUNK]) :].reshape(-1)[: int(blob[PK.H_N_CHUNKS]) * PK.CHUNK_WORDS].reshape(-1, PK.CHUNK_WORDS)
    S = [0, 0, 0, 0]
    re = np.float32(0)
    im = np.float32(0)
    if int(lvl[0]) == 0:
        return ("approx", np.float32(0), np.float32(0))
    for c in range(int(lvl[7]), int(lvl[7]) + int(lvl[8])):
        off, _, ng, _ = (int(v) for v in chunks[c])
        for _g in range(ng):
            h = [int(v) for v in data[off : off + FAST_HEADER_WORDS]]
            nL, nPi, nD = h[0] & 0xFFF, (h[0] >> 12) & 0xFFF, h[0] >> 24
            a = h[1]
            o = off + FAST_HEADER_WORDS
            for j in range(nL):
                r = data[o : o + SL]
                a = (a + _par(xw, r[:W]) * int(r[W])) & M32
                o += SL
            o = off + FAST_HEADER_WORDS + round4(nL * SL)
            e = 0
            for j in range(nPi):
                r = data[o : o + SP]
                e ^= _par(xw, r[:W]) & _par(xw, r[W : 2 * W])
                o += SP
            a = (a + (e << 31)) & M32
            o = off + FAST_HEADER_WORDS + round4(nL * SL) + round4(nPi * SP)
            SM = mpair_stride(W)
            for j in range(h[7]):
                r = data[o : o + SM]
                pa, pb = _par(xw, r[:W]), _par(xw, r[W : 2 * W])
                a = (a + pa * int(r[2 * W]) + pb * int(r[2 * W + 1]) + (pa & pb) * int(r[2 * W + 2])) & M32
                o += SM
            Pp = (1, 0, 0, 0)
            for j in range(nD):
                r = data[o : o + SD]
                pa, pb = _par(xw, r[:W]), _par(xw, r[W : 2 * W])
                f = PAIR[(int(r[2 * W]) ^ (pa << 2) ^ (pb << 5)) & 63]
                Pp = f if j == 0 else _mul(Pp, f)
                o += SD
            if (a & 0xFFFF) == 0:
                Pb, Qb = PELL[(a >> 16) & 127]
                k1, k2 = h[8:12], h[12:16]
                v = tuple((k1[i] * Pb + k2[i] * Qb) & M32 for i in range(4))
                v = _rotate(v, a >> 29)
                if nD:
                    v = _mul(v, Pp)
                if not approx:
                    sc = 1 << h[4]
                    S = [(S[i] + v[i] * sc) & M32 for i in range(4)]
                else:
                    vc = np.array([_s32(t) for t in v], dtype=np.int32)
                    tre, tim = to_complex_parts(vc[None, :], np.array([0], np.int32))
                    sc = np.array([h[2]], np.uint32).view(np.float32)[0]
                    with np.errstate(all="ignore"):
                        tre, tim = (tre * sc).astype(np.float32), (tim * sc).astype(np.float32)
                    are = np.array([h[5]], np.uint32).view(np.float32)[0]
                    aim = np.array([h[6]], np.uint32).view(np.float32)[0]
                    with np.errstate(all="ignore"):
                        ure = np.float32(np.float32(tre[0] * are) - np.float32(tim[0] * aim))
                        uim = np.float32(np.float32(tre[0] * aim) + np.float32(tim[0] * are))
                        pw = np.array([h[3]], np.uint32).view(np.float32)[0]
                        re = np.float32(re + np.float32(ure * pw))
                        im = np.float32(im + np.float32(uim * pw))
            off = o
    if approx:
        return ("approx", re, im)
    c = [_s32(v) for v in S]
    p = p_lo
    t = (c[0] | c[1] | c[2] | c[3]) & M32
    if t:
        sh = (t & -t).bit_length() - 1
        c = [v >> sh for v in c]
        p += sh
    return ("exact", np.array(c, np.int32), p)
