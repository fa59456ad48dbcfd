// MODE_SLICED pipeline: bit-sliced evaluation, 32 shots per thread.
//
//   K0t transpose_in_kernel   packed f rows            -> XT[param row][slab]   (one 32-bit word = 32 shots)
//   K1s sample_sliced_kernel  XT + g (TMA-staged)      -> OT[draw][slab]        (bit-sliced output bits)
//   K2a assemble_out_kernel   OT + direct bits of f    -> packed output rows
//   K1c norm_check_kernel     shot 0 re-evaluated with the per-row evaluator (needs the companion fast/faithful blob)
//
// K1s: a thread owns a slab of 32 shots.  Its parameter matrix lives transposed in shared memory (a private column
// per thread: row i = bit i of the 32 shots), so the GF(2) contraction of a term for all 32 shots is the XOR of the
// rows its mask selects -- no popcount, cost proportional to the mask weight.  The exponents of the monoid element
// w^a (1+sqrt2)^b are accumulated as bit-planes; only the decode of a graph's value, the sum over graphs and the
// draw run per shot (fully unrolled over the slab, accumulators in registers).  Record format: pack_sliced.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blob.h"
#include "sampler_kernels.cuh"
#include "zomega.cuh"

namespace tsb {

constexpr int kSlicedMaxThreads = 256;
constexpr int kSlicedHeaderWords = 20;
constexpr int kMaxGeneralPairs = 8;

struct SParams {
  const uint32_t* __restrict__ blob;     // sliced blob
  const uint32_t* __restrict__ xt;       // [sum_c F_c][slab_cap]
  uint32_t* __restrict__ ot;             // [n_draws][slab_cap]
  const uint32_t* __restrict__ subkeys;  // [n_draws][2]
  long long B;
  long long shot_offset;
  int n_slabs;
  int slab_cap;
  int per_cta;  // slabs per CTA per round
  int rounds;
  int resident;
  int n_stages;
  int stage_words;
  int smem_xt_off;    // word offsets inside dynamic shared memory
  int smem_pw_off;
  int smem_s_off;
  int smem_prev_off;
  int smem_data_off;
  int rows;           // zero_row + 1
};

struct SlicedTables {
  int4 pair[64];
  int2 pell[128];
};

// ---------------------------------------------------------------------------------------------
// K0t: f rows -> transposed words.  One warp per slab, lane = shot.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_in_kernel(const uint32_t* __restrict__ blob, const uint64_t* __restrict__ f,
                                                           long long B, int n_slabs, int slab_cap, uint32_t* __restrict__ xt) {
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= n_slabs) return;
  const long long row = (long long)warp * 32 + lane;
  const bool active = row < B;
  const int wf = (int)blob[H_WF64];
  const uint32_t* __restrict__ fsel = blob + blob[H_OFF_FSEL];
  const uint32_t* __restrict__ comp_tab = blob + blob[H_OFF_COMP];
  const int n_comp = (int)blob[H_N_COMP];
  int total = 0;
  if (n_comp > 0) {
    const uint32_t* last = comp_tab + (n_comp - 1) * kCompWords;
    total = (int)(last[C_FSEL_OFF] + last[C_F]);
  }
  for (int i0 = 0; i0 < total; i0 += 32) {
    uint32_t mine = 0;
    const int lim = min(32, total - i0);
    for (int j = 0; j < lim; ++j) {
      const uint32_t fi = fsel[i0 + j];
      const uint32_t bit = active ? (uint32_t)((f[row * wf + (fi >> 6)] >> (fi & 63u)) & 1ull) : 0u;
      const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
      if (lane == j) mine = word;
    }
    if (lane < lim) xt[(size_t)(i0 + lane) * slab_cap + warp] = mine;
  }
}

// ---------------------------------------------------------------------------------------------
// K2a: bit-sliced outputs + direct bits -> packed rows.  One warp per slab, lane = shot.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assemble_out_kernel(const uint32_t* __restrict__ blob, const uint64_t* __restrict__ f,
                                                           const uint32_t* __restrict__ ot, long long B, int n_slabs, int slab_cap,
                                                           uint64_t* __restrict__ out) {
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= n_slabs) return;
  const long long row = (long long)warp * 32 + lane;
  if (row >= B) return;
  const int wf = (int)blob[H_WF64], wo = (int)blob[H_WOUT64];
  const int n_direct = (int)blob[H_N_DIRECT], n_draws = (int)blob[H_N_DRAWS];
  const uint32_t* __restrict__ direct_tab = blob + blob[H_OFF_DIRECT];
  const uint32_t* __restrict__ dest = blob + blob[H_OFF_DEST];
  for (int w = 0; w < wo; ++w) {
    uint64_t v = 0;
    for (int j = 0; j < n_direct; ++j) {
      const uint32_t fi = direct_tab[2 * j], dd = direct_tab[2 * j + 1];
      const uint32_t d = dd & 0x7FFFFFFFu;
      if ((int)(d >> 6) != w) continue;
      const uint64_t bit = ((f[row * wf + (fi >> 6)] >> (fi & 63u)) & 1ull) ^ (uint64_t)(dd >> 31);
      v |= bit << (d & 63u);
    }
    for (int j = 0; j < n_draws; ++j) {
      const uint32_t d = dest[j];
      if ((int)(d >> 6) != w) continue;
      const uint64_t bit = (ot[(size_t)j * slab_cap + warp] >> lane) & 1u;
      v |= bit << (d & 63u);
    }
    out[row * wo + w] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// K1s
// ---------------------------------------------------------------------------------------------
// Per-thread state lives in private shared-memory columns (index [row][tid], so a warp access is one conflict-free
// wavefront): the transposed parameters xt, the parity words of general pairs pw, the per-shot level accumulators S and
// the chain-rule state prev.  Keeping S / prev out of registers lets the per-shot loops stay rolled: the hot loop
// body must fit the 32 KB instruction cache (a 32x unrolled body ran at 26 % issue, stalled on instruction fetch).
__device__ __forceinline__ void add_a3(uint32_t& A0, uint32_t& A1, uint32_t& A2, uint32_t da, uint32_t p) {
  if (da & 1u) {
    const uint32_t c0 = A0 & p;
    A0 ^= p;
    const uint32_t c1 = A1 & c0;
    A1 ^= c0;
    A2 ^= c1;
  }
  if (da & 2u) {
    const uint32_t c1 = A1 & p;
    A1 ^= p;
    A2 ^= c1;
  }
  if (da & 4u) A2 ^= p;
}

__device__ __forceinline__ void add_cnt5(uint32_t (&Bp)[5], uint32_t w) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const uint32_t t = Bp[k] & w;
    Bp[k] ^= w;
    w = t;
  }
}

// XOR of the rows named by index words; row r of this thread's column sits at xcol[r * T].
template <int T>
__device__ __forceinline__ uint32_t ld_row(const uint32_t* __restrict__ xcol, uint32_t row) {
  return xcol[row * (uint32_t)T];
}

template <int T>
__device__ __forceinline__ uint32_t xor4(const uint32_t* __restrict__ xcol, uint32_t iw) {
  const uint32_t a0 = ld_row<T>(xcol, iw & 255u), a1 = ld_row<T>(xcol, __byte_perm(iw, 0, 0x4441));
  const uint32_t a2 = ld_row<T>(xcol, __byte_perm(iw, 0, 0x4442)), a3 = ld_row<T>(xcol, iw >> 24);
  return (a0 ^ a1) ^ (a2 ^ a3);
}

template <int T>
__device__ __forceinline__ void rows4(const uint32_t* __restrict__ xcol, uint32_t iw, uint32_t (&r)[4]) {
  r[0] = ld_row<T>(xcol, iw & 255u);
  r[1] = ld_row<T>(xcol, __byte_perm(iw, 0, 0x4441));
  r[2] = ld_row<T>(xcol, __byte_perm(iw, 0, 0x4442));
  r[3] = ld_row<T>(xcol, iw >> 24);
}

template <int T>
__device__ __forceinline__ uint32_t xor12(const uint32_t* __restrict__ xcol, uint32_t i0, uint32_t i1, uint32_t i2) {
  uint32_t a[4], b[4], c[4];
  rows4<T>(xcol, i0, a);
  rows4<T>(xcol, i1, b);
  rows4<T>(xcol, i2, c);
  return (a[0] ^ a[1] ^ a[2]) ^ (a[3] ^ b[0] ^ b[1]) ^ (b[2] ^ b[3] ^ c[0]) ^ (c[1] ^ c[2] ^ c[3]);
}

template <int T>
__device__ __forceinline__ void xor24(const uint32_t* __restrict__ xcol, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t j0, uint32_t j1,
                                      uint32_t j2, uint32_t& p1, uint32_t& p2) {
  uint32_t a[4], b[4], c[4], d[4], e[4], f[4];
  rows4<T>(xcol, i0, a);
  rows4<T>(xcol, i1, b);
  rows4<T>(xcol, i2, c);
  rows4<T>(xcol, j0, d);
  rows4<T>(xcol, j1, e);
  rows4<T>(xcol, j2, f);
  p1 = (a[0] ^ a[1] ^ a[2]) ^ (a[3] ^ b[0] ^ b[1]) ^ (b[2] ^ b[3] ^ c[0]) ^ (c[1] ^ c[2] ^ c[3]);
  p2 = (d[0] ^ d[1] ^ d[2]) ^ (d[3] ^ e[0] ^ e[1]) ^ (e[2] ^ e[3] ^ f[0]) ^ (f[1] ^ f[2] ^ f[3]);
}

template <int T>
__device__ __forceinline__ uint32_t sliced_parity(const uint32_t* __restrict__ sdata, uint32_t o, int n, const uint32_t* __restrict__ xcol) {
  uint32_t acc0 = 0, acc1 = 0;
  int w = 0;
  for (; w + 1 < n; w += 2) {
    const uint32_t i0 = sdata[o + w], i1 = sdata[o + w + 1];
    acc0 ^= xor4<T>(xcol, i0);
    acc1 ^= xor4<T>(xcol, i1);
  }
  if (w < n) acc0 ^= xor4<T>(xcol, sdata[o + w]);
  return acc0 ^ acc1;
}

template <int T, bool HAS_EXACT>
__device__ __forceinline__ void sliced_graphs(const uint32_t* __restrict__ sdata, uint32_t off, int n_graphs, bool approx,
                                              const uint32_t* __restrict__ xcol, uint32_t* __restrict__ pwcol, uint32_t* __restrict__ scol,
                                              const SlicedTables* __restrict__ tb) {
  for (int g = 0; g < n_graphs; ++g) {
    const uint4 h0 = *reinterpret_cast<const uint4*>(sdata + off);
    const uint4 h1 = *reinterpret_cast<const uint4*>(sdata + off + 4);
    const int n_terms = (int)(h0.x & 0xFFFFu), n_gen = (int)(h0.x >> 16);
    const uint32_t b64 = h0.y & 0xFFu;
    uint32_t A0 = 0, A1 = 0, A2 = 0, Z = 0;
    uint32_t Bp[5] = {0, 0, 0, 0, 0};
    // ---- phase 1: bit-sliced accumulation over the term stream (software-pipelined: the next term's record is
    //      fetched while the current term's rows are in flight)
    uint32_t o = off + kSlicedHeaderWords;
    uint4 cur0 = *reinterpret_cast<const uint4*>(sdata + o), cur1 = *reinterpret_cast<const uint4*>(sdata + o + 4);
    for (int t = 0; t < n_terms; ++t) {
      const uint32_t cw = cur0.x;
      const uint32_t type = cw & 3u;
      const int n1 = (int)((cw >> 2) & 63u), n2 = (int)((cw >> 8) & 63u);
      const bool generic = (cw >> 31) != 0u;
      uint32_t len;
      if (!generic) len = type == 0u ? 4u : 8u;
      else len = (uint32_t)round4(1 + (type == 3u ? 1 : 0) + n1 + (type != 0u ? n2 : 0));
      const uint4 nx0 = *reinterpret_cast<const uint4*>(sdata + o + len);
      const uint4 nx1 = *reinterpret_cast<const uint4*>(sdata + o + len + 4);
      uint32_t p1, p2 = 0, ex = 0;
      if (!generic) {
        // compact record: twelve rows per parity, loaded unconditionally (padding names the zero row) so that all
        // loads of the term are in flight before the first XOR needs one
        if (type == 0u) {
          p1 = xor12<T>(xcol, cur0.y, cur0.z, cur0.w);
        } else {
          xor24<T>(xcol, cur0.y, cur0.z, cur0.w, cur1.x, cur1.y, cur1.z, p1, p2);
          ex = cur1.w;
        }
      } else {
        const uint32_t o1 = o + 1 + (type == 3u ? 1u : 0u);
        if (type == 3u) ex = sdata[o + 1];
        p1 = sliced_parity<T>(sdata, o1, n1, xcol);
        if (type != 0u) p2 = sliced_parity<T>(sdata, o1 + n1, n2, xcol);
      }
      if (type == 0u) {
        add_a3(A0, A1, A2, (cw >> 14) & 7u, p1);
        const uint32_t bm = (cw >> 17) & 3u, zm = (cw >> 19) & 3u;
        if (bm) add_cnt5(Bp, bm == 1u ? p1 : ~p1);
        if (zm) Z |= (zm == 1u ? p1 : ~p1);
      } else if (type == 1u) {
        A2 ^= p1 & p2;
      } else if (type == 2u) {
        const uint32_t slot = (cw >> 14) & 15u;
        pwcol[(2 * slot) * T] = p1;
        pwcol[(2 * slot + 1) * T] = p2;
      } else {
        const uint32_t wd[3] = {p1, p2, p1 & p2};
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          add_a3(A0, A1, A2, (ex >> (6 * v)) & 7u, wd[v]);
          const int db = (int)((ex >> (6 * v + 3)) & 7u) - 3;
          const uint32_t w = db > 0 ? wd[v] : ~wd[v];
          for (int r = 0; r < (db < 0 ? -db : db); ++r) add_cnt5(Bp, w);
        }
        const uint32_t ztt = (ex >> 18) & 15u;
        if (ztt & 1u) Z |= ~p1 & ~p2;
        if (ztt & 2u) Z |= p1 & ~p2;
        if (ztt & 4u) Z |= ~p1 & p2;
        if (ztt & 8u) Z |= p1 & p2;
      }
      o += len;
      cur0 = nx0;
      cur1 = nx1;
    }
    // ---- phase 2: per-shot decode and accumulation
    const uint4 k1 = *reinterpret_cast<const uint4*>(sdata + off + 8);
    const uint4 k2 = *reinterpret_cast<const uint4*>(sdata + off + 12);
    const uint2 ctlw = *reinterpret_cast<const uint2*>(sdata + off + 16);
    const float are = __uint_as_float(h1.y), aim = __uint_as_float(h1.z);
    const float sc = pow2_f32((int)h0.z), pw = pow2_f32((int)h0.w);
    const uint32_t fx = 1u << (h1.x & 31u);
    // cnt: five count planes -> one word per plane; a: three planes
#pragma unroll 4
    for (int s = 0; s < 32; ++s) {
      if ((Z >> s) & 1u) continue;
      const uint32_t a = ((A0 >> s) & 1u) | (((A1 >> s) & 1u) << 1) | (((A2 >> s) & 1u) << 2);
      const uint32_t cnt = ((Bp[0] >> s) & 1u) | (((Bp[1] >> s) & 1u) << 1) | (((Bp[2] >> s) & 1u) << 2) |
                           (((Bp[3] >> s) & 1u) << 3) | (((Bp[4] >> s) & 1u) << 4);
      const int2 pq = tb->pell[(b64 + cnt) & 127u];
      const uint32_t P = (uint32_t)pq.x, Q = (uint32_t)pq.y;
      ZW v = ZW{k1.x * P + k2.x * Q, k1.y * P + k2.y * Q, k1.z * P + k2.z * Q, k1.w * P + k2.w * Q};
      v = zw_rotate(v, a);
      for (int slot = 0; slot < n_gen; ++slot) {
        const uint32_t ctl = ((slot < 4 ? ctlw.x : ctlw.y) >> (8 * (slot & 3))) & 63u;
        const uint32_t pa = (pwcol[(2 * slot) * T] >> s) & 1u;
        const uint32_t pb = (pwcol[(2 * slot + 1) * T] >> s) & 1u;
        v = zw_mul(v, zw_from(tb->pair[(ctl ^ (pa << 2) ^ (pb << 5)) & 63u]));
      }
      if (!approx) {
        if constexpr (HAS_EXACT) {
          uint4* sp = reinterpret_cast<uint4*>(scol) + s * T;
          uint4 acc = *sp;
          acc.x += v.c0 * fx; acc.y += v.c1 * fx; acc.z += v.c2 * fx; acc.w += v.c3 * fx;
          *sp = acc;
        }
      } else {
        const float s2 = TSB_SQRT1_2;
        const float f0 = __int2float_rn((int32_t)v.c0), f1 = __int2float_rn((int32_t)v.c1);
        const float f2 = __int2float_rn((int32_t)v.c2), f3 = __int2float_rn((int32_t)v.c3);
        const float t1 = __fmul_rn(f1, s2), t3 = __fmul_rn(f3, s2);
        const float tre = __fmul_rn(__fadd_rn(__fadd_rn(f0, t1), t3), sc);
        const float tim = __fmul_rn(__fsub_rn(__fadd_rn(t1, f2), t3), sc);
        const float ure = __fsub_rn(__fmul_rn(tre, are), __fmul_rn(tim, aim));
        const float uim = __fadd_rn(__fmul_rn(tre, aim), __fmul_rn(tim, are));
        float2* sp = reinterpret_cast<float2*>(scol) + s * (HAS_EXACT ? 2 : 1) * T;
        float2 acc = *sp;
        acc.x = __fadd_rn(acc.x, __fmul_rn(ure, pw));
        acc.y = __fadd_rn(acc.y, __fmul_rn(uim, pw));
        *sp = acc;
      }
    }
    off += h1.w;
  }
}

// dynamic shared memory (32-bit words): [0,64) mbarriers | SlicedTables | xt [rows][T] | pw [16][T] |
// S [32][T] x (uint4 if HAS_EXACT else float2) | prev [32][T] | data region / stage ring
template <int T, bool HAS_EXACT>
__global__ void __launch_bounds__(T, 1) sample_sliced_kernel(const SParams prm) {
  extern __shared__ __align__(128) uint32_t smem[];
  const uint32_t* __restrict__ blob = prm.blob;
  const int tid = threadIdx.x;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  SlicedTables* tb = reinterpret_cast<SlicedTables*>(smem + 64);
  uint32_t* xcol = smem + prm.smem_xt_off + tid;  // this thread's column: row r at xcol[r * T]
  uint32_t* pwcol = smem + prm.smem_pw_off + tid;
  // S: element (s, tid) at scol + s * stride * T words, stride = 4 (exact) or 2 (approx) words; vector accesses
  uint32_t* scol = smem + prm.smem_s_off + tid * (HAS_EXACT ? 4 : 2);
  float* pcol = reinterpret_cast<float*>(smem + prm.smem_prev_off) + tid;
  uint32_t* sdata = smem + prm.smem_data_off;

  const int n_comp = (int)blob[H_N_COMP], n_chunks = (int)blob[H_N_CHUNKS];
  const uint32_t* __restrict__ comp_tab = blob + blob[H_OFF_COMP];
  const uint32_t* __restrict__ level_tab = blob + blob[H_OFF_LEVEL];
  const uint32_t* __restrict__ chunk_tab = blob + blob[H_OFF_CHUNK];
  const uint32_t* __restrict__ gdata = blob + blob[H_OFF_DATA];
  const int one_row = (int)blob[H_ONE_ROW];

  for (int i = tid; i < 64; i += T) {
    int a = i & 7, b = i >> 3;
    int4 ua = unit_phase(a), ub = unit_phase(b), uc = unit_phase(a + b);
    tb->pair[i] = make_int4(1 + ua.x + ub.x - uc.x, ua.y + ub.y - uc.y, ua.z + ub.z - uc.z, ua.w + ub.w - uc.w);
  }
  if (tid == 0) {
    uint32_t P = 1, Q = 0;
    for (int e = 0; e < 64; ++e) { tb->pell[64 + e] = make_int2((int)P, (int)Q); uint32_t nP = P + 2u * Q, nQ = P + Q; P = nP; Q = nQ; }
    P = 1; Q = 0;
    for (int e = 0; e <= 64; ++e) { tb->pell[64 - e] = make_int2((int)P, (int)Q); uint32_t nP = 2u * Q - P, nQ = P - Q; P = nP; Q = nQ; }
  }
  const int n_bars = prm.resident ? 1 : prm.n_stages;
  if (tid == 0) {
    for (int i = 0; i < n_bars; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();

  const long long total_q = (long long)prm.rounds * n_chunks;
  auto issue = [&](long long q) {
    const int ch = (int)(q % n_chunks);
    const uint32_t* row = chunk_tab + ch * kChunkWords;
    const int stage = (int)(q % prm.n_stages);
    const uint32_t bytes = row[K_WORDS] * 4u;
    mbar_expect_tx(&bars[stage], bytes);
    tma_bulk_g2s(sdata + (size_t)stage * prm.stage_words, gdata + row[K_OFF], bytes, &bars[stage]);
  };
  if (tid == 0 && n_chunks > 0) {
    if (prm.resident) {
      mbar_expect_tx(&bars[0], blob[H_DATA_WORDS] * 4u);
      for (int ch = 0; ch < n_chunks; ++ch) {
        const uint32_t* row = chunk_tab + ch * kChunkWords;
        tma_bulk_g2s(sdata + row[K_OFF], gdata + row[K_OFF], row[K_WORDS] * 4u, &bars[0]);
      }
    } else {
      for (long long q = 0; q < prm.n_stages && q < total_q; ++q) issue(q);
    }
  }
  if (prm.resident && n_chunks > 0) mbar_wait(&bars[0], 0);

  long long q = 0;
  for (int round = 0; round < prm.rounds; ++round) {
    const int slab = (round * (int)gridDim.x + (int)blockIdx.x) * prm.per_cta + tid;
    const bool active = tid < prm.per_cta && slab < prm.n_slabs;
    // a warp with no slab at all only keeps the stage ring's barriers company
    const int wslab = (round * (int)gridDim.x + (int)blockIdx.x) * prm.per_cta + (tid & ~31);
    const bool warp_active = (tid & ~31) < prm.per_cta && wslab < prm.n_slabs;
    const unsigned long long shot0 = (unsigned long long)prm.shot_offset + (unsigned long long)slab * 32ull;

    int xt_row0 = 0, draw0 = 0;
    for (int ci = 0; ci < n_comp; ++ci) {
      const uint32_t* __restrict__ comp = comp_tab + ci * kCompWords;
      const int F = (int)comp[C_F], n_c = (int)comp[C_NC];
      for (int i = 0; i < F; ++i) xcol[i * T] = active ? prm.xt[(size_t)(xt_row0 + i) * prm.slab_cap + slab] : 0u;
      for (int i = F; i < prm.rows; ++i) xcol[i * T] = 0u;
      xcol[one_row * T] = 0xFFFFFFFFu;

      for (int k = 0; k <= n_c; ++k) {
        const uint32_t* __restrict__ lvl = level_tab + (comp[C_FIRST_LEVEL] + k) * kLevelWords;
        const bool approx = (lvl[L_FLAGS] & 1u) != 0u;
        if (k > 0) xcol[(F + k - 1) * T] = 0xFFFFFFFFu;  // trying bit 1 for every shot
        for (int s = 0; s < 32; ++s) {
          if constexpr (HAS_EXACT) reinterpret_cast<uint4*>(scol)[s * T] = make_uint4(0u, 0u, 0u, 0u);
          else reinterpret_cast<uint2*>(scol)[s * T] = make_uint2(0u, 0u);
        }
        const int first_chunk = (int)lvl[L_FIRST_CHUNK], nck = (int)lvl[L_N_CHUNKS];
        for (int c = 0; c < nck; ++c) {
          const uint32_t* row = chunk_tab + (first_chunk + c) * kChunkWords;
          uint32_t off;
          if (prm.resident) {
            off = row[K_OFF];
          } else {
            const int stage = (int)(q % prm.n_stages);
            mbar_wait(&bars[stage], (uint32_t)((q / prm.n_stages) & 1));
            off = (uint32_t)stage * (uint32_t)prm.stage_words;
          }
          if (warp_active) sliced_graphs<T, HAS_EXACT>(sdata, off, (int)row[K_GRAPHS], approx, xcol, pwcol, scol, tb);
          if (!prm.resident) {
            __syncthreads();
            if (tid == 0 && q + prm.n_stages < total_q) issue(q + prm.n_stages);
            ++q;
          }
        }
        // finish the level for the 32 shots: |amp|, draw, chain rule
        const int p_lo = (int)lvl[L_P_LO];
        const bool empty = lvl[L_G] == 0u;
        uint32_t k0 = 0, k1 = 0;
        if (k > 0) {
          k0 = prm.subkeys[2 * (draw0 + k - 1)];
          k1 = prm.subkeys[2 * (draw0 + k - 1) + 1];
        }
        uint32_t bits = 0;
#pragma unroll 2
        for (int s = 0; s < (warp_active ? 32 : 0); ++s) {
          float re, im;
          if (approx) {
            const float2 a2 = reinterpret_cast<const float2*>(scol)[s * (HAS_EXACT ? 2 : 1) * T];
            re = a2.x; im = a2.y;
          } else {
            if constexpr (HAS_EXACT) {
              const uint4 a4 = reinterpret_cast<const uint4*>(scol)[s * T];
              ZW cz = ZW{a4.x, a4.y, a4.z, a4.w};
              int p = p_lo;
              zw_fixpoint(cz, p);
              zw_to_complex(cz, p, re, im);
            } else {
              re = 0.0f; im = 0.0f;
            }
          }
          if (empty) { re = 0.0f; im = 0.0f; }
          const float p1 = complex_abs(re, im);
          if (k == 0) {
            pcol[s * T] = p1;
          } else {
            const float pv = pcol[s * T];
            const float u = uniform_f32(k0, k1, shot0 + (unsigned long long)s);
            const bool bit = u < __fdiv_rn(p1, pv);
            pcol[s * T] = bit ? p1 : __fsub_rn(pv, p1);
            bits |= (bit ? 1u : 0u) << s;
          }
        }
        if (k > 0) {
          xcol[(F + k - 1) * T] = bits;
          if (active) prm.ot[(size_t)(draw0 + k - 1) * prm.slab_cap + slab] = bits;
        }
      }
      xt_row0 += F;
      draw0 += n_c;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1c: normalisation check of in-batch shot 0 (sampler.py:66-72) with the per-row evaluator on the companion blob.
// One CTA per component; evaluation t = 0 is level 0, t = 2k-1 is level k with trying bit 1, t = 2k with bit 0.
// Dynamic shared memory: (2 n_c + 1) floats.
// ---------------------------------------------------------------------------------------------
template <int W, int MODE>
__global__ void __launch_bounds__(128) norm_check_kernel(const uint32_t* __restrict__ blob, const uint64_t* __restrict__ f_row0,
                                                         const uint64_t* __restrict__ out_row0, float* __restrict__ norm_dev) {
  __shared__ Tables tb;
  extern __shared__ float vals[];
  init_tables(&tb, threadIdx.x, blockDim.x);
  __syncthreads();
  const int ci = blockIdx.x;
  const uint32_t* __restrict__ comp = blob + blob[H_OFF_COMP] + ci * kCompWords;
  const int F = (int)comp[C_F], n_c = (int)comp[C_NC], first_draw = (int)comp[C_FIRST_DRAW];
  const uint32_t* __restrict__ sel = blob + blob[H_OFF_FSEL] + comp[C_FSEL_OFF];
  const uint32_t* __restrict__ dest = blob + blob[H_OFF_DEST];
  const uint32_t* __restrict__ chunk_tab = blob + blob[H_OFF_CHUNK];
  const GmemSrc src{blob + blob[H_OFF_DATA]};
  const int n_evals = 2 * n_c + 1;
  for (int t = threadIdx.x; t < n_evals; t += blockDim.x) {
    const int k = (t + 1) >> 1;
    const uint32_t trybit = (uint32_t)(t & 1);
    uint32_t x[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      uint32_t xv = 0;
      const int lim = min(32, F - 32 * w);
      for (int b = 0; b < lim; ++b) {
        const uint32_t fi = sel[32 * w + b];
        xv |= (uint32_t)((f_row0[fi >> 6] >> (fi & 63u)) & 1ull) << b;
      }
      for (int j = 0; j < k; ++j) {
        const int pos = F + j;
        if ((pos >> 5) != w) continue;
        uint32_t bit;
        if (j == k - 1) {
          bit = trybit;
        } else {
          const uint32_t d = dest[first_draw + j];
          bit = (uint32_t)((out_row0[d >> 6] >> (d & 63u)) & 1ull);
        }
        xv |= bit << (pos & 31);
      }
      x[w] = xv;
    }
    if (MODE == kModeFast) x[W - 1] |= 0x80000000u;
    const uint32_t* __restrict__ lvl = blob + blob[H_OFF_LEVEL] + (comp[C_FIRST_LEVEL] + k) * kLevelWords;
    LevelAcc acc;
    acc.reset();
    const int first_chunk = (int)lvl[L_FIRST_CHUNK], nck = (int)lvl[L_N_CHUNKS];
    for (int c = 0; c < nck; ++c) {
      const uint32_t* row = chunk_tab + (first_chunk + c) * kChunkWords;
      eval_chunk<W, MODE>(src, row[K_OFF], (int)row[K_GRAPHS], lvl, x, acc, &tb);
    }
    float re, im;
    finish_level<MODE>(acc, (lvl[L_FLAGS] & 1u) != 0u, (int)lvl[L_P_LO], re, im);
    if (lvl[L_G] == 0u) { re = 0.0f; im = 0.0f; }
    vals[t] = complex_abs(re, im);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float prev = vals[0], dev = 0.0f;
    for (int k = 1; k <= n_c; ++k) {
      const float p1 = vals[2 * k - 1], p0 = vals[2 * k];
      const float norm = __fdiv_rn(__fadd_rn(p0, p1), prev);
      const float d = fabsf(__fsub_rn(norm, 1.0f));
      dev = (dev != dev || d != d) ? __uint_as_float(0x7FC00000u) : fmaxf(dev, d);
      const uint32_t dd = dest[first_draw + k - 1];
      const bool bit = ((out_row0[dd >> 6] >> (dd & 63u)) & 1ull) != 0ull;
      prev = bit ? p1 : __fsub_rn(prev, p1);
    }
    norm_dev[ci] = dev;
  }
}

}  // namespace tsb
