"""Pure-Python model of the MODE_SLICED device evaluation (bit-sliced over an arbitrary number of shots).

Words are Python integers with one bit per shot, so the model follows the kernel's plane arithmetic literally
(XOR of rows per parity, 3-plane adder for ``a``, ripple counter for ``b``, OR-plane for vanishing factors) and
then, per shot, gathers the plane bits into an index and adds the decode-table entry exactly like the kernel's
second phase.  CPU test infrastructure only.
"""

import numpy as np

from fast_model import M32, _s32
from tsim_b200 import pack as PK
from tsim_b200.pack_sliced import CLASS_WORDS, PAIR_TABLE, PI_CLASSES, SLICED_HEADER_WORDS, _mul_u32


def evaluate_level(pp: PK.PackedProgram, comp: int, level: int, x_bits: np.ndarray):
    """x_bits: uint8 [N, P].  -> list over shots of ("exact", coeffs, power) / ("approx", re, im)."""
    blob = pp.blob
    assert pp.mode == PK.MODE_SLICED
    N, P = x_bits.shape
    full = (1 << N) - 1
    one_row, zero_row = int(blob[PK.H_ONE_ROW]), int(blob[PK.H_ZERO_ROW])
    rows = {i: int("".join("1" if x_bits[s, i] else "0" for s in reversed(range(N))) or "0", 2) for i in range(P)}
    rows[one_row] = full
    rows[zero_row] = 0
    comp_row = blob[int(blob[PK.H_OFF_COMP]) + comp * PK.COMP_WORDS :]
    lrow = int(comp_row[4]) + level
    lvl = blob[int(blob[PK.H_OFF_LEVEL]) + lrow * PK.LEVEL_WORDS :][: PK.LEVEL_WORDS]
    approx = bool(lvl[6] & 1)
    p_lo = _s32(int(lvl[9]))
    data = blob[int(blob[PK.H_OFF_DATA]) :]
    chunks = blob[int(blob[PK.H_OFF_CHUNK]) :][: int(blob[PK.H_N_CHUNKS]) * PK.CHUNK_WORDS].reshape(-1, PK.CHUNK_WORDS)
    S = [[0, 0, 0, 0] for _ in range(N)]
    RE = [np.float32(0)] * N
    IM = [np.float32(0)] * N
    if int(lvl[0]) == 0:
        return [("approx", np.float32(0), np.float32(0))] * N

    scale = int(blob[PK.H_INDEX_SCALE])  # index bytes hold row * scale
    assert scale in (1, 2)

    def row(byte):
        assert byte % scale == 0
        return rows.get(byte // scale, 0)

    def par4(w):
        return row(w & 255) ^ row((w >> 8) & 255) ^ row((w >> 16) & 255) ^ row(w >> 24)

    for c in range(int(lvl[7]), int(lvl[7]) + int(lvl[8])):
        coff, _, ng, _ = (int(v) for v in chunks[c])
        for _g in range(ng):
            off = coff + int(data[coff + _g])  # directory: chunk-relative record offsets
            h = [int(v) for v in data[off : off + SLICED_HEADER_WORDS]]
            n_words, n_gen = h[0] & 0xFFFF, h[0] >> 16
            n_idx, nb, n_mul = h[1] & 0xFF, (h[1] >> 8) & 0xFF, (h[1] >> 16) & 0xFF
            mul_ctl = [(int(data[off + (h[3] & 0xFFFF) - 4]) >> (8 * j)) & 63 for j in range(n_mul)]
            main_words = h[3] >> 16
            assert 0 < main_words <= n_words or n_words == 0
            tbl = coff + h[2]
            A = [0, 0, 0]
            Bp = [0] * 5
            Z = 0
            gen = {}

            def add_a(da, p):
                if da & 1:
                    c0 = A[0] & p
                    A[0] ^= p
                    c1 = A[1] & c0
                    A[1] ^= c0
                    A[2] ^= c1
                if da & 2:
                    c1 = A[1] & p
                    A[1] ^= p
                    A[2] ^= c1
                if da & 4:
                    A[2] ^= p

            def add_cnt(wd):
                cy = wd
                for k in range(nb):
                    t = Bp[k] & cy
                    Bp[k] ^= cy
                    cy = t

            def lin_op(prm, p):
                nonlocal Z
                add_a(prm & 7, p)
                bm, zm = (prm >> 3) & 3, (prm >> 5) & 3
                if bm == 1:
                    add_cnt(p)
                elif bm == 2:
                    add_cnt(~p & full)
                if zm == 1:
                    Z |= p
                elif zm == 2:
                    Z |= ~p & full

            def pair_op(op, prm, pa, pb):
                nonlocal Z
                if op == 3:
                    gen[prm & 15] = (pa, pb)
                    return
                assert op == 4
                ex = prm
                for v, wd in enumerate((pa, pb, pa & pb)):
                    add_a((ex >> (6 * v)) & 7, wd)
                    db = ((ex >> (6 * v + 3)) & 7) - 3
                    for _ in range(abs(db)):
                        add_cnt(wd if db > 0 else (~wd & full))
                ztt = (ex >> 18) & 15
                for combo in range(4):
                    if (ztt >> combo) & 1:
                        wa = pa if combo & 1 else ~pa & full
                        wb = pb if combo & 2 else ~pb & full
                        Z |= wa & wb

            def par_words(o, n):
                p = 0
                for w in data[o : o + n]:
                    p ^= par4(int(w))
                return p

            o = off + SLICED_HEADER_WORDS
            end = o + n_words
            aux_at = o + main_words  # [main | aux]: a helper warp walks the aux part into a plane of its own
            saved = None
            while o < end:
                if o == aux_at and saved is None:
                    saved = (list(A), list(Bp), Z, dict(gen))
                    A[:] = [0, 0, 0]
                    Bp[:] = [0] * 5
                    Z = 0
                    gen.clear()
                kind, count = int(data[o]) & 0xFFFF, int(data[o]) >> 16
                o += 4
                if saved is not None:
                    assert 3 <= kind < 9 or 18 <= kind < 21 or kind == 22  # pi runs only
                if kind in (16, 17):  # LIN_1 / LIN2_1: [params, index word]
                    for _i in range(count):
                        prm = int(data[o])
                        if kind == 17:
                            assert prm in (0, 2)  # 0: the alignment filler over the all-zeros row
                            prm = 2
                        lin_op(prm, par_words(o + 1, 1))
                        o += 2
                    continue
                if 18 <= kind < 21:  # PI_1 + k: [index word, 3 index words]
                    for _i in range(count):
                        A[2] ^= par_words(o, 1) & par_words(o + 1, kind - 17)
                        o += 4
                    continue
                if kind == 21:  # PAIR_1: [op | params << 3, index word, index word, 0]
                    for _i in range(count):
                        hdr = int(data[o])
                        pair_op(hdr & 7, hdr >> 3, par_words(o + 1, 1), par_words(o + 2, 1))
                        o += 4
                    continue
                if kind < 3 or 9 <= kind < 12:  # LIN / LIN2 runs
                    cls = kind % 3
                    nw = CLASS_WORDS[cls]
                    for _i in range(count):
                        if kind >= 9:
                            assert int(data[o]) == 2
                        lin_op(int(data[o]), par_words(o + 1, nw))
                        o += 4 if cls < 2 else 8
                    continue
                if 12 <= kind < 15:  # PAIR runs
                    nw = CLASS_WORDS[kind - 12]
                    for _i in range(count):
                        hdr = int(data[o])
                        pair_op(hdr & 7, hdr >> 3, par_words(o + 4, nw), par_words(o + 8, nw))
                        o += 12
                    continue
                if kind < 9:  # PI runs
                    c1, c2 = PI_CLASSES[kind - 3]
                    for _i in range(count):
                        A[2] ^= par_words(o, CLASS_WORDS[c1]) & par_words(o + 4, CLASS_WORDS[c2])
                        o += 8
                    continue
                assert kind in (15, 22)  # generic block stream (22: pi terms only)
                gend = o + count
                q = 0
                while o < gend:
                    hdr = int(data[o])
                    op, prm = hdr & 7, hdr >> 3
                    nw = int(data[o + 1])
                    p = par_words(o + 2, nw)
                    o += 2 + nw
                    if op == 0:
                        q = p
                    elif op == 1:
                        lin_op(prm, p)
                    elif op == 2:
                        A[2] ^= q & p
                    else:
                        pair_op(op, prm, q, p)
                assert o == gend
            assert o == end
            if saved is not None:
                # the aux part may only have XORed into the top plane of a: that is what lets phase 2 merge it with one XOR
                assert A[0] == 0 and A[1] == 0 and not any(Bp) and Z == 0 and not gen
                aux_a2 = A[2]
                A[:], Bp[:], Z = saved[0], saved[1], saved[2]
                gen.update(saved[3])
                A[2] ^= aux_a2
            for s in range(N):
                if (Z >> s) & 1:
                    continue
                planes = A + Bp[:nb]
                for slot in range(n_gen):
                    planes = planes + list(gen[slot])
                assert len(planes) == n_idx
                idx = sum(((pl >> s) & 1) << k for k, pl in enumerate(planes))
                if not approx:
                    e = [int(v) for v in data[tbl + 4 * idx : tbl + 4 * idx + 4]]
                    for j, ctl in enumerate(mul_ctl):  # two-stage decode: ring factors of the multiplied general pairs
                        pa, pb = ((gen[j][0] >> s) & 1), ((gen[j][1] >> s) & 1)
                        fac = PAIR_TABLE[ctl ^ (pa << 2) ^ (pb << 5)]
                        e = [int(v) for v in _mul_u32(np.array(e, dtype=np.uint32), fac)]
                    S[s] = [(S[s][i] + e[i]) & M32 for i in range(4)]
                else:
                    e = data[tbl + 2 * idx : tbl + 2 * idx + 2].view(np.float32)
                    with np.errstate(all="ignore"):
                        RE[s] = np.float32(RE[s] + e[0])
                        IM[s] = np.float32(IM[s] + e[1])
    out = []
    for s in range(N):
        if approx:
            out.append(("approx", RE[s], IM[s]))
            continue
        cs = [_s32(v) for v in S[s]]
        p = p_lo
        t = (cs[0] | cs[1] | cs[2] | cs[3]) & M32
        if t:
            sh = (t & -t).bit_length() - 1
            cs = [v >> sh for v in cs]
            p += sh
        out.append(("exact", np.array(cs, np.int32), p))
    return out
