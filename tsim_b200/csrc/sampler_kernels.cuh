// K1: fused evaluate + autoregressive sample kernel, and K3: evaluate-only kernel.
//
// Work mapping (both modes): one shot per thread.  The packed g_{tki} records of a graph are the same
// for every shot, so every shared-memory read in the inner loops is a warp-uniform broadcast; the
// shot's parameter vector (its selected f bits followed by the outputs drawn so far) lives in W
// registers; a term's GF(2) contraction is W AND/XORs and one popcount.  All CTAs are persistent
// (grid = #SMs x CTAs/SM) and walk shot tiles; the record stream is staged into shared memory with
// TMA bulk copies (cp.async.bulk + mbarrier) -- once per CTA when all of g fits ("resident"), else
// chunk by chunk through a ring of stages.
//
// MODE_FAITHFUL evaluates exactly the reference's sequence of int32 operations
// (terms.py:56-187, evaluate.py:33-59, exact_scalar.py:98-137).  MODE_FAST is used only when
// pack.py proved that no int32 operation of the reference can wrap for this program; it replaces the
// node/half-pi/pi families by one packed-integer accumulation in the monoid
// {w^a (1+sqrt2)^b (1+w)^n} and sums graphs in fixed point -- same canonical (coeffs, power).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blob.h"
#include "zomega.cuh"

namespace tsb {

// threads per CTA of the per-row sampling kernel: more resident warps hide the POPC / LDS latencies (measured on cfg2:
// 512 -> 3.82 ms, 768 -> 3.27 ms, 1024 -> 3.27 ms); wide parameter vectors need the registers of the 512-thread shape
__host__ __device__ constexpr int threads_for_words(int W) { return W <= 4 ? 768 : 512; }

struct KParams {
  const uint32_t* __restrict__ blob;  // whole blob in HBM
  const uint64_t* __restrict__ f;     // [B, wf64] packed error-mechanism rows
  uint64_t* __restrict__ out;         // [B, wout64] packed output rows
  float* __restrict__ norm_dev;       // [n_components]
  const uint32_t* __restrict__ subkeys;  // [n_draws][2]
  long long B;
  long long shot_offset;
  int n_tiles;
  int resident;     // 1: data region copied to smem once
  int n_stages;     // streaming ring depth
  int stage_words;  // words per stage
  int smem_data_off;  // word offset of the data region / stage ring inside dynamic smem
  // optional indirection (pattern-cache pass 2): process rows rows[0 .. *n_rows) instead of 0 .. B-1
  const uint32_t* __restrict__ rows;
  const uint32_t* __restrict__ n_rows;
};

// ---------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (sm_90+ PTX; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// record sources: shared memory (K1) or global memory (K3)
// ---------------------------------------------------------------------------------------------
struct SmemSrc {
  const uint32_t* base;  // dynamic shared memory
  __device__ __forceinline__ uint32_t ld(uint32_t off) const { return base[off]; }
  __device__ __forceinline__ uint2 ld2(uint32_t off) const { return *reinterpret_cast<const uint2*>(base + off); }
  __device__ __forceinline__ uint4 ld4(uint32_t off) const { return *reinterpret_cast<const uint4*>(base + off); }
};
struct GmemSrc {
  const uint32_t* base;
  __device__ __forceinline__ uint32_t ld(uint32_t off) const { return __ldg(base + off); }
  __device__ __forceinline__ uint2 ld2(uint32_t off) const { return __ldg(reinterpret_cast<const uint2*>(base + off)); }
  __device__ __forceinline__ uint4 ld4(uint32_t off) const { return __ldg(reinterpret_cast<const uint4*>(base + off)); }
};

// small per-CTA lookup tables (shared memory)
struct Tables {
  int4 one_plus[8];  // 1 + w^k                       terms.py:37
  int4 unit[8];      // w^k                           terms.py:22-34
  int4 pair[64];     // 1 + w^a + w^b - w^(a+b), index a | b << 3     terms.py:179-181
  int2 pell[128];    // (1+sqrt2)^(b-64) = P + Q sqrt2, index b
};

__device__ __forceinline__ int4 unit_phase(int k) {
  switch (k & 7) {
    case 0: return make_int4(1, 0, 0, 0);
    case 1: return make_int4(0, 1, 0, 0);
    case 2: return make_int4(0, 0, 1, 0);
    case 3: return make_int4(0, 0, 0, -1);
    case 4: return make_int4(-1, 0, 0, 0);
    case 5: return make_int4(0, -1, 0, 0);
    case 6: return make_int4(0, 0, -1, 0);
    default: return make_int4(0, 0, 0, 1);
  }
}

__device__ inline void init_tables(Tables* t, int tid, int nthreads) {
  for (int i = tid; i < 8; i += nthreads) {
    int4 u = unit_phase(i);
    t->unit[i] = u;
    u.x += 1;
    t->one_plus[i] = u;
  }
  for (int i = tid; i < 64; i += nthreads) {
    int a = i & 7, b = i >> 3;
    int4 ua = unit_phase(a), ub = unit_phase(b), uc = unit_phase(a + b);
    t->pair[i] = make_int4(1 + ua.x + ub.x - uc.x, ua.y + ub.y - uc.y, ua.z + ub.z - uc.z, ua.w + ub.w - uc.w);
  }
  if (tid == 0) {
    // (1+sqrt2)^e = P + Q sqrt2: (P,Q) -> (P+2Q, P+Q); inverse (sqrt2-1): (P,Q) -> (2Q-P, P-Q). Wraps silently
    // for |e| > 24; such entries are unreachable when pack.py's bound holds.
    uint32_t P = 1, Q = 0;
    for (int e = 0; e < 64; ++e) {
      t->pell[64 + e] = make_int2((int)P, (int)Q);
      uint32_t nP = P + 2u * Q, nQ = P + Q;
      P = nP; Q = nQ;
    }
    P = 1; Q = 0;
    for (int e = 0; e <= 64; ++e) {
      t->pell[64 - e] = make_int2((int)P, (int)Q);
      uint32_t nP = 2u * Q - P, nQ = P - Q;
      P = nP; Q = nQ;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-level accumulator (one evaluate() call for one parameter vector)
// ---------------------------------------------------------------------------------------------
struct LevelAcc {
  ZW c;       // exact running sum
  int p;
  int started;
  float re, im;  // approximate branch
  __device__ __forceinline__ void reset() {
    c = zw_make(0, 0, 0, 0); p = 0; started = 0; re = 0.0f; im = 0.0f;
  }
};

template <int W>
__device__ __forceinline__ int parity(const uint32_t (&x)[W], const uint32_t* m) {
  uint32_t t = x[0] & m[0];
#pragma unroll
  for (int w = 1; w < W; ++w) t ^= x[w] & m[w];
  return __popc(t) & 1;
}

template <int W, class Src>
__device__ __forceinline__ int parity_src(const uint32_t (&x)[W], const Src& src, uint32_t off) {
  uint32_t t = x[0] & src.ld(off);
#pragma unroll
  for (int w = 1; w < W; ++w) t ^= x[w] & src.ld(off + w);
  return __popc(t) & 1;
}

// ---- MODE_FAITHFUL --------------------------------------------------------------------------------
// One chunk = n_graphs records laid out as in pack.py::_faithful_level_records.
template <int W, class Src>
__device__ __forceinline__ void eval_chunk_faithful(const Src& src, uint32_t off, int n_graphs, int A, int H, int C, int D,
                                                    bool approx, const uint32_t (&x)[W], LevelAcc& acc, const Tables* tb) {
  for (int g = 0; g < n_graphs; ++g) {
    // node phases: fold(mul_with_power) over all A slots, padded slots multiply by the identity (terms.py:66-73)
    ZW N = zw_make(1, 0, 0, 0);
    int Np = 0;
    for (int j = 0; j < A; ++j) {
      int par = parity_src<W>(x, src, off);
      uint32_t ctl = src.ld(off + W);
      off += W + 1;
      int4 f = (ctl & 8u) ? tb->one_plus[((par << 2) + (int)ctl) & 7] : make_int4(1, 0, 0, 0);
      if (j == 0) {
        N = zw_from(f);
      } else {
        N = zw_mul(N, zw_from(f));
        zw_reduce1(N, Np);
      }
    }
    if (A > 0) zw_fixpoint(N, Np);
    // half-pi phases (terms.py:104-107)
    int h = 0;
    for (int j = 0; j < H; ++j) {
      int par = parity_src<W>(x, src, off);
      h += par * (int)(src.ld(off + W) & 7u);
      off += W + 1;
    }
    // pi products (terms.py:136-144)
    int e = 0;
    for (int j = 0; j < C; ++j) {
      int p1 = parity_src<W>(x, src, off);
      int p2 = parity_src<W>(x, src, off + W);
      uint32_t cst = src.ld(off + 2 * W);
      off += 2 * W + 1;
      e ^= (p1 ^ (int)(cst & 1u)) & (p2 ^ (int)((cst >> 1) & 1u));
    }
    // phase pairs (terms.py:174-187)
    ZW Pp = zw_make(1, 0, 0, 0);
    int Pq = 0;
    for (int j = 0; j < D; ++j) {
      int pa = parity_src<W>(x, src, off);
      int pb = parity_src<W>(x, src, off + W);
      uint32_t ctl = src.ld(off + 2 * W);
      off += 2 * W + 1;
      int a = ((int)(ctl & 7u) + 4 * pa) & 7, b = ((int)((ctl >> 3) & 7u) + 4 * pb) & 7;
      int4 f = (ctl & 64u) ? tb->pair[a | (b << 3)] : make_int4(1, 0, 0, 0);
      if (j == 0) {
        Pp = zw_from(f);
      } else {
        Pp = zw_mul(Pp, zw_from(f));
        zw_reduce1(Pp, Pq);
      }
    }
    if (D > 0) zw_fixpoint(Pp, Pq);
    // prefactor + six-way plain product (evaluate.py:37-50); wrapping products commute, unit phases merge
    uint32_t phase = src.ld(off);
    ZW ff = zw_make((int)src.ld(off + 1), (int)src.ld(off + 2), (int)src.ld(off + 3), (int)src.ld(off + 4));
    int power2 = (int)src.ld(off + 5);
    float are = __uint_as_float(src.ld(off + 6)), aim = __uint_as_float(src.ld(off + 7));
    off += kPrefactorWords;
    ZW T = zw_mul(N, zw_from(tb->unit[(h + 4 * e + (int)phase) & 7]));
    T = zw_mul(T, Pp);
    T = zw_mul(T, ff);
    int Tp = Np + Pq;
    if (!approx) {
      // evaluate.py:52-54: fold(add_with_power) over graphs in order
      if (!acc.started) {
        acc.c = T; acc.p = Tp + power2; acc.started = 1;
      } else {
        zw_add_p(acc.c, acc.p, T, Tp + power2);
      }
    } else {
      // evaluate.py:56-59 with the op order of oracle/evaluation.py
      float tre, tim;
      zw_to_complex(T, Tp, tre, tim);
      float ure = __fsub_rn(__fmul_rn(tre, are), __fmul_rn(tim, aim));
      float uim = __fadd_rn(__fmul_rn(tre, aim), __fmul_rn(tim, are));
      float pw = pow2_f32(power2);
      acc.re = __fadd_rn(acc.re, __fmul_rn(ure, pw));
      acc.im = __fadd_rn(acc.im, __fmul_rn(uim, pw));
    }
  }
}

// ---- MODE_FAST ------------------------------------------------------------------------------------
// Record strides (words); keep in sync with tsim_b200/pack_fast.py
__host__ __device__ constexpr int round4(int n) { return (n + 3) & ~3; }
__host__ __device__ constexpr int fast_lin_stride(int W) { return W == 1 ? 2 : round4(W + 1); }
__host__ __device__ constexpr int fast_pi_stride(int W) { return W == 1 ? 2 : round4(2 * W); }
__host__ __device__ constexpr int fast_pair_stride(int W) { return round4(2 * W + 1); }
__host__ __device__ constexpr int fast_mpair_stride(int W) { return round4(2 * W + 3); }
constexpr int kFastHeaderWords = 16;
// graph header: [0] nL | nPi << 12 | nD << 24   [1] acc0   [2] float32 bits of 2^p_T (p_T = n >> 2)   [3] float32 bits of 2^power2
//               [4] shift (= p_T + power2 - p_lo)   [5] approx.re  [6] approx.im  [7] number of monoid pairs
//               [8..11] K1 = (1+w)^(n&3) * floatfactor   [12..15] K2 = K1 * sqrt2
// acc fields:   bits 0..15 count of vanishing node factors, bits 16..28 b + 64, bits 29..31 a

template <int N, class Src>
__device__ __forceinline__ void load_rec(const Src& src, uint32_t off, uint32_t (&r)[N]) {
  if constexpr (N == 2) {
    uint2 v = src.ld2(off);
    r[0] = v.x; r[1] = v.y;
  } else {
    static_assert(N % 4 == 0, "record strides are 2 or a multiple of 4 words");
#pragma unroll
    for (int w = 0; w < N; w += 4) {
      uint4 v = src.ld4(off + w);
      r[w] = v.x; r[w + 1] = v.y; r[w + 2] = v.z; r[w + 3] = v.w;
    }
  }
}

template <int W>
__device__ __forceinline__ uint32_t masked_xor(const uint32_t (&x)[W], const uint32_t* m) {
  uint32_t t = x[0] & m[0];
#pragma unroll
  for (int w = 1; w < W; ++w) t ^= x[w] & m[w];
  return t;
}

__device__ __forceinline__ ZW zw_rotate(ZW v, uint32_t a) {
  // multiply by w^a: w*(c0,c1,c2,c3) = (c3, c0, c1, -c2); i*(..) = (-c2, c3, c0, -c1)
  if (a & 1u) v = ZW{v.c3, v.c0, v.c1, 0u - v.c2};
  if (a & 2u) v = ZW{0u - v.c2, v.c3, v.c0, 0u - v.c1};
  if (a & 4u) v = ZW{0u - v.c0, 0u - v.c1, 0u - v.c2, 0u - v.c3};
  return v;
}

template <int W, class Src>
__device__ __forceinline__ void eval_chunk_fast(const Src& src, uint32_t off, int n_graphs, bool approx,
                                                const uint32_t (&x)[W], LevelAcc& acc, const Tables* tb) {
  constexpr int SL = fast_lin_stride(W), SP = fast_pi_stride(W), SD = fast_pair_stride(W);
  for (int g = 0; g < n_graphs; ++g) {
    const uint4 h0 = src.ld4(off);
    const int nL = h0.x & 0xFFF, nPi = (h0.x >> 12) & 0xFFF, nD = h0.x >> 24;
    uint32_t a = h0.y;
    uint32_t o = off + kFastHeaderWords;
    // linear terms (node + half-pi families): a += parity(x & mask) * delta
    for (int j = 0; j < nL; ++j, o += SL) {
      uint32_t r[SL];
      load_rec<SL>(src, o, r);
      // a += parity * delta as one IMAD (the compiler's own choice is compare + select + add)
      asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a) : "r"((uint32_t)__popc(masked_xor<W>(x, r)) & 1u), "r"(r[W]));
    }
    o = off + kFastHeaderWords + round4(nL * SL);
    // pi terms: a += 4 * (parity(psi) & parity(phi)); the constants ride on the always-one parameter bit
    uint32_t e = 0;
    for (int j = 0; j < nPi; ++j, o += SP) {
      uint32_t r[SP];
      load_rec<SP>(src, o, r);
      e ^= (uint32_t)(__popc(masked_xor<W>(x, r)) & __popc(masked_xor<W>(x, r + W)));
    }
    a += e << 31;
    o = off + kFastHeaderWords + round4(nL * SL) + round4(nPi * SP);
    // monoid-type pairs: a += pa * d1 + pb * d2 + (pa & pb) * d12
    {
      constexpr int SM = fast_mpair_stride(W);
      const int nM = (int)src.ld(off + 7);
      for (int j = 0; j < nM; ++j, o += SM) {
        uint32_t r[SM];
        load_rec<SM>(src, o, r);
        const uint32_t pa = (uint32_t)__popc(masked_xor<W>(x, r)) & 1u, pb = (uint32_t)__popc(masked_xor<W>(x, r + W)) & 1u;
        a += pa * r[2 * W] + pb * r[2 * W + 1] + (pa & pb) * r[2 * W + 2];
      }
    }
    // general pair terms: plain (wrapping) product of table factors
    ZW Pp = zw_make(1, 0, 0, 0);
    for (int j = 0; j < nD; ++j, o += SD) {
      uint32_t r[SD];
      load_rec<SD>(src, o, r);
      uint32_t pa = (uint32_t)__popc(masked_xor<W>(x, r)) & 1u, pb = (uint32_t)__popc(masked_xor<W>(x, r + W)) & 1u;
      ZW f = zw_from(tb->pair[(r[2 * W] ^ (pa << 2) ^ (pb << 5)) & 63u]);
      Pp = (j == 0) ? f : zw_mul(Pp, f);
    }
    // decode: value = 2^pT * w^a * (1+sqrt2)^b * K1
    if ((a & 0xFFFFu) == 0u) {  // no vanishing node factor
      const uint4 h1 = src.ld4(off + 4), k1 = src.ld4(off + 8), k2 = src.ld4(off + 12);
      const int2 pq = tb->pell[(a >> 16) & 127u];
      const uint32_t P = (uint32_t)pq.x, Q = (uint32_t)pq.y;
      ZW v = ZW{k1.x * P + k2.x * Q, k1.y * P + k2.y * Q, k1.z * P + k2.z * Q, k1.w * P + k2.w * Q};
      v = zw_rotate(v, a >> 29);
      if (nD) v = zw_mul(v, Pp);
      if (!approx) {
        const uint32_t sc = 1u << h1.x;  // shift < 31 is guaranteed by pack.py's bound
        acc.c.c0 += v.c0 * sc; acc.c.c1 += v.c1 * sc; acc.c.c2 += v.c2 * sc; acc.c.c3 += v.c3 * sc;
      } else {
        // zw_to_complex with the scale 2^p_T read from the record
        const float s2 = TSB_SQRT1_2;
        const float f0 = __int2float_rn((int32_t)v.c0), f1 = __int2float_rn((int32_t)v.c1);
        const float f2 = __int2float_rn((int32_t)v.c2), f3 = __int2float_rn((int32_t)v.c3);
        const float t1 = __fmul_rn(f1, s2), t3 = __fmul_rn(f3, s2), sc = __uint_as_float(h0.z);
        const float tre = __fmul_rn(__fadd_rn(__fadd_rn(f0, t1), t3), sc);
        const float tim = __fmul_rn(__fsub_rn(__fadd_rn(t1, f2), t3), sc);
        const float are = __uint_as_float(h1.y), aim = __uint_as_float(h1.z);
        float ure = __fsub_rn(__fmul_rn(tre, are), __fmul_rn(tim, aim));
        float uim = __fadd_rn(__fmul_rn(tre, aim), __fmul_rn(tim, are));
        float pw = __uint_as_float(h0.w);
        acc.re = __fadd_rn(acc.re, __fmul_rn(ure, pw));
        acc.im = __fadd_rn(acc.im, __fmul_rn(uim, pw));
      }
    }
    off = o;
  }
}

// |evaluate(level)| from a finished accumulator.  p_lo: common power of the fixed-point sum (fast mode).
template <int MODE>
__device__ __forceinline__ void finish_level(const LevelAcc& acc, bool approx, int p_lo, float& re, float& im) {
  if (approx) {
    re = acc.re; im = acc.im;
    return;
  }
  ZW c = acc.c;
  int p = (MODE == kModeFast) ? p_lo : acc.p;
  zw_fixpoint(c, p);
  zw_to_complex(c, p, re, im);
}

template <int W, int MODE, class Src>
__device__ __forceinline__ void eval_chunk(const Src& src, uint32_t off, int n_graphs, const uint32_t* __restrict__ lvl,
                                           const uint32_t (&x)[W], LevelAcc& acc, const Tables* tb) {
  const bool approx = (lvl[L_FLAGS] & 1u) != 0u;
  if constexpr (MODE == kModeFaithful) {
    eval_chunk_faithful<W>(src, off, n_graphs, (int)lvl[L_A], (int)lvl[L_H], (int)lvl[L_C], (int)lvl[L_D], approx, x, acc, tb);
  } else {
    eval_chunk_fast<W>(src, off, n_graphs, approx, x, acc, tb);
  }
}

}  // namespace tsb
