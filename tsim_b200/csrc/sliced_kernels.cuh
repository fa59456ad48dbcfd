// MODE_SLICED pipeline: the default sampling path, bit-sliced over shots.
//
//   K0t transpose_in_kernel     packed f rows            -> XT[param row][slab]   (one 32-bit word = 32 shots)
//   K1s sample_sliced_kernel    XT + g (TMA-staged)      -> OT[draw][slab]        (bit-sliced output bits)
//   K2a assemble_out_kernel     OT + direct bits of f    -> packed output rows
//   K1c norm_check_fast_kernel  shot 0 re-evaluated on the companion per-row blob (norm_check_kernel: sequential fallback)
//
// What K1s replaces in the reference (all file:line relative to /root/reference/src/tsim):
//   * the chain rule over the outputs of a component, `_sample_component` (sampler.py:45-81): level k evaluates
//     |E_k(f, m_<k, 1)|, draws bit k with jax.random.bernoulli on the in-batch shot index (threefry, sampler.py:74-75),
//     keeps prev = bit ? p1 : prev - p1; components in program order with the key threaded through (sampler.py:134-148);
//   * `evaluate` (compile/evaluate.py:33-59): per graph, the product of the four term families (compile/terms.py:56-73
//     node phases, :94-107 half-pi phases, :125-144 pi products, :164-187 phase pairs) times the prefactor, summed over
//     graphs -- exactly (Z[w] fixed point, core/exact_scalar.py:52-137) or, with approximate float factors, as a
//     sequential complex64 sum in graph order (evaluate.py:56-59);
//   * `matmul_gf2` (utils/linalg.py:81-102): every parity <mask, x> mod 2.
// Here a parity for 32 shots is the XOR of the rows of the transposed parameter matrix that the mask names (no GEMM, no
// popcount); the families' monoid exponents (pack_fast.py) accumulate as bit-planes; a graph's value times its prefactor
// is looked up in a decode table built at pack time (pack_sliced.py) with the reference's float32 operation order, so
// sums, |amp| and draws see the same bits as the per-row kernels and the oracle.  Record format: pack_sliced.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "blob.h"
#include "sampler_kernels.cuh"
#include "zomega.cuh"

namespace tsb {

constexpr int kSlicedHeaderWords = 8;
constexpr int kMinPlaneRows = 12;  // "vanished" plane + up to 11 index planes (pack_sliced.MAX_INDEX_BITS); more with multiplied pairs

struct SParams {
  const uint32_t* __restrict__ blob;     // sliced blob
  const uint32_t* __restrict__ xt;       // [sum_c F_c][slab_cap]
  uint32_t* __restrict__ ot;             // [n_draws][slab_cap]
  float* __restrict__ pv;                // [slab_cap][32]: chain-rule state |amplitude of the prefix| per shot
  const uint32_t* __restrict__ subkeys;  // [n_draws][2]
  long long B;
  long long shot_offset;
  int n_slabs;
  int slab_cap;
  int n_groups;  // units of 32 slabs (1024 shots): narrow groups, or half the lanes of a wide group
  int ng;        // units per CTA per round
  int rounds;
  int n_stages;
  int stage_words;
  int smem_xt_off;  // word offsets inside dynamic shared memory
  int smem_pl_off;
  int smem_data_off;
  int rows;  // zero_row + 1
  int plane_rows;  // plane rows per graph slot (blob header H_PLANE_ROWS)
  uint4 sel;    // dp4a byte selectors {s << 0, s << 8, s << 16, s << 24}, s = row stride in bytes / H_INDEX_SCALE (see par4)
  uint4 sel_e;  // the same with 8, the byte size of a float2 decode-table entry, instead of 128 (see sliced_phase2)
  // optional indirection (HAS_ROWS kernels; memoised path pass 2, post-selection survivors): slot i of the launch is
  // batch row rows[i], i < *n_rows -- the count is only known on the device, so the kernel derives n_slabs, n_groups and
  // rounds itself.  The row index is also the shot's RNG counter (sampler.py:75: bernoulli over the whole batch).
  const uint32_t* __restrict__ row_list;
  const uint32_t* __restrict__ n_rows;
  // narrow layout: the transposes of K0t / K2a done by the groups themselves (no separate launches, no xt round trip):
  // f_rows != nullptr -> a group builds its matrix from the packed f rows; out_rows != nullptr -> it assembles its shots'
  // packed output rows (direct bits + drawn bits) when it is through with the last component
  const uint64_t* __restrict__ f_rows;
  uint64_t* __restrict__ out_rows;
  // ring flow control: 0 = warps count themselves out of a stage and the last one refills it (groups drift apart: one
  // group's phase 2 overlaps another's phase 1); 1 = one CTA-wide barrier per chunk (all groups in the same code at the
  // same time: fewer instruction-cache misses, which wins for the larger exact-level kernels -- cfg4: 13.0 vs 11.8 ms)
  int lockstep;
  // component / level / chunk tables staged in shared memory behind the ring when there is room (thin launches: the
  // walk through the program is latency-bound and these reads sit on its critical path); -1: read them through L2
  int smem_tab_off;
  int tab_words;
};

// ---------------------------------------------------------------------------------------------
// K0t: f rows -> transposed words.  One warp per slab, lane = shot; a block of eight slabs writes each row as one sector.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_in_kernel(const uint32_t* __restrict__ blob, const uint64_t* __restrict__ f,
                                                           long long B, int n_slabs, int slab_cap, uint32_t* __restrict__ xt,
                                                           const uint32_t* __restrict__ row_list, const uint32_t* __restrict__ n_rows) {
  __shared__ uint32_t tile[32][9];  // [row][slab of the block]: rows leave as 32-byte bursts of eight slabs
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab0 = (int)blockIdx.x * 8, warp = slab0 + wl;
  if (row_list) {  // the launch covers the worst case; the live count sits on the device
    B = (long long)*n_rows;
    n_slabs = (int)((B + 31) / 32);
    if (slab0 >= n_slabs) return;
  }
  const long long slot = (long long)warp * 32 + lane;
  const bool active = warp < n_slabs && slot < B;
  const long long row = (row_list && active) ? (long long)row_list[slot] : slot;
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
    tile[lane][wl] = mine;
    __syncthreads();
    const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
    if (r < lim && slab0 + c < n_slabs) xt[(size_t)(i0 + r) * slab_cap + slab0 + c] = tile[r][c];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// K2a: bit-sliced outputs + direct bits -> packed rows.  One warp per slab, lane = shot.  (The sampling kernel assembles
// its own rows in the narrow layout; this kernel serves the wide layout, rows wider than two words and programs without
// compiled components, e.g. rank-1 circuits whose outputs are all direct.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assemble_out_kernel(const uint32_t* __restrict__ blob, const uint64_t* __restrict__ f,
                                                           const uint32_t* __restrict__ ot, long long B, int n_slabs, int slab_cap,
                                                           uint64_t* __restrict__ out, const uint32_t* __restrict__ row_list,
                                                           const uint32_t* __restrict__ n_rows) {
  __shared__ uint32_t s_tab[2 * 256];  // a tile of the direct table (f index, destination | flip << 31)
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row_list) {
    B = (long long)*n_rows;
    n_slabs = (int)((B + 31) / 32);
  }
  const long long slot = (long long)warp * 32 + lane;
  const bool valid = warp < n_slabs && slot < B;
  const long long row = valid ? (row_list ? (long long)row_list[slot] : slot) : 0;
  const int wf = (int)blob[H_WF64], wo = (int)blob[H_WOUT64];
  const int n_direct = (int)blob[H_N_DIRECT], n_draws = (int)blob[H_N_DRAWS];
  const uint32_t* __restrict__ direct_tab = blob + blob[H_OFF_DIRECT];
  const uint32_t* __restrict__ dest = blob + blob[H_OFF_DEST];
  if (wf <= 4 && wo <= 2) {
    // one pass over the direct table (staged in shared memory tile by tile), the shot's f words in registers
    uint64_t fw0 = 0, fw1 = 0, fw2 = 0, fw3 = 0;
    if (valid) {
      const uint64_t* fr = f + row * wf;
      fw0 = fr[0];
      if (wf > 1) fw1 = fr[1];
      if (wf > 2) fw2 = fr[2];
      if (wf > 3) fw3 = fr[3];
    }
    uint64_t v0 = 0, v1 = 0;
    for (int j0 = 0; j0 < n_direct; j0 += 256) {
      const int lim = min(256, n_direct - j0);
      __syncthreads();
      if ((int)threadIdx.x < lim) {
        s_tab[2 * threadIdx.x] = direct_tab[2 * (j0 + threadIdx.x)];
        s_tab[2 * threadIdx.x + 1] = direct_tab[2 * (j0 + threadIdx.x) + 1];
      }
      __syncthreads();
#pragma unroll 4
      for (int j = 0; j < lim; ++j) {
        const uint32_t fi = s_tab[2 * j], dd = s_tab[2 * j + 1];
        const uint32_t ws = fi >> 6;
        const uint64_t fw = ws == 0 ? fw0 : ws == 1 ? fw1 : ws == 2 ? fw2 : fw3;
        const uint64_t bit = ((fw >> (fi & 63u)) & 1ull) ^ (uint64_t)(dd >> 31);
        if (dd & 64u) v1 |= bit << (dd & 63u);  // warp-uniform: destination word 1
        else v0 |= bit << (dd & 63u);
      }
    }
    if (valid) {
      for (int j = 0; j < n_draws; ++j) {
        const uint32_t d = dest[j];
        const uint64_t bit = (ot[(size_t)j * slab_cap + warp] >> lane) & 1u;
        if (d & 64u) v1 |= bit << (d & 63u);
        else v0 |= bit << (d & 63u);
      }
      out[row * wo] = v0;
      if (wo > 1) out[row * wo + 1] = v1;
    }
    return;
  }
  if (!valid) return;
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
// A *group* of SPLIT warps owns 32 lanes of slabs.  In the narrow layout a lane is one slab (32 shots, one 32-bit word
// per parameter row; the group holds 1024 shots); in the wide layout a lane is two consecutive slabs (64 shots, one
// 64-bit word per row, LDS.64; 2048 shots per group), which halves the instructions per shot of phase 1 -- the row
// address, the record fetch and the loop control are shared by twice as many shots.  The group's transposed parameter
// matrix sits in shared memory ([row][lane], conflict-free wavefronts).  The graphs of a chunk are handed out in waves
// of SPLIT: warp w of the group runs phase 1 (bit-sliced term stream -> planes) for graph wave + w and parks the
// planes in shared memory; after a group barrier every warp runs phase 2 for ITS 32 / SPLIT shots of each slab over
// the wave's graphs in order (the approximate branch is a sequential float sum over graphs): gather the plane bits
// into the index, add the decode-table entry.  Level accumulators stay in registers.
//
// Wide groups are filled in *units* of 16 lanes (32 slabs = 1024 shots, the same granularity as a narrow group): when
// a CTA holds an odd number of units, lanes 16..31 of its last group sit out: 64-bit shared-memory accesses are served
// per half-warp, so an idle half costs no wavefront and shared-memory traffic stays proportional to the shots.
struct U2 {
  uint32_t x, y;
};
__device__ __forceinline__ U2 operator^(U2 a, U2 b) { return U2{a.x ^ b.x, a.y ^ b.y}; }
__device__ __forceinline__ U2 operator&(U2 a, U2 b) { return U2{a.x & b.x, a.y & b.y}; }
__device__ __forceinline__ U2 operator|(U2 a, U2 b) { return U2{a.x | b.x, a.y | b.y}; }
__device__ __forceinline__ U2 operator~(U2 a) { return U2{~a.x, ~a.y}; }
__device__ __forceinline__ U2& operator^=(U2& a, U2 b) { a.x ^= b.x; a.y ^= b.y; return a; }
__device__ __forceinline__ U2& operator|=(U2& a, U2 b) { a.x |= b.x; a.y |= b.y; return a; }

template <class LW>
struct LaneWord;
template <>
struct LaneWord<uint32_t> {
  static constexpr int kWords = 1;
  static __device__ __forceinline__ uint32_t zero() { return 0u; }
  static __device__ __forceinline__ uint32_t ones() { return 0xFFFFFFFFu; }
  static __device__ __forceinline__ uint32_t ld(uint32_t saddr) {
    return *reinterpret_cast<const uint32_t*>(__cvta_shared_to_generic(saddr));
  }
};
template <>
struct LaneWord<U2> {
  static constexpr int kWords = 2;
  static __device__ __forceinline__ U2 zero() { return U2{0u, 0u}; }
  static __device__ __forceinline__ U2 ones() { return U2{0xFFFFFFFFu, 0xFFFFFFFFu}; }
  static __device__ __forceinline__ U2 ld(uint32_t saddr) {
    const uint2 v = *reinterpret_cast<const uint2*>(__cvta_shared_to_generic(saddr));
    return U2{v.x, v.y};
  }
};

template <class LW>
__device__ __forceinline__ void add_a3(LW& A0, LW& A1, LW& A2, uint32_t da, LW p) {
  if (da & 1u) {
    const LW c0 = A0 & p;
    A0 ^= p;
    const LW c1 = A1 & c0;
    A1 ^= c0;
    A2 ^= c1;
  }
  if (da & 2u) {
    const LW c1 = A1 & p;
    A1 ^= p;
    A2 ^= c1;
  }
  if (da & 4u) A2 ^= p;
}

template <class LW>
__device__ __forceinline__ void add_cnt5(LW (&Bp)[5], LW w) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const LW t = Bp[k] & w;
    Bp[k] ^= w;
    w = t;
  }
}

// Row indices in the term stream are bytes, four per word.  xs = 32-bit shared-memory address of this lane's column of
// the group's matrix (row r at xs + stride r; stride = 128 bytes narrow, 256 wide).  One dp4a with a selector word
// (s in byte k, 0 elsewhere) turns byte k into the row's address on the FMA pipe, which leaves the ALU pipe to the
// XORs.  The selectors arrive as kernel parameters so that they stay constant-bank operands; index bytes are stored
// times H_INDEX_SCALE (2 when the program has at most 127 rows: s = 64 narrow, 128 wide; else 1: s = 128, narrow only).
template <class LW>
__device__ __forceinline__ LW par4(uint32_t xs, uint32_t w, const uint4& sel) {
  typedef LaneWord<LW> L;
  return (L::ld(__dp4a(w, sel.x, xs)) ^ L::ld(__dp4a(w, sel.y, xs))) ^ (L::ld(__dp4a(w, sel.z, xs)) ^ L::ld(__dp4a(w, sel.w, xs)));
}
template <class LW, int N>
__device__ __forceinline__ LW par_words(uint32_t xs, const uint32_t* w, const uint4& sel) {
  LW p = par4<LW>(xs, w[0], sel);
#pragma unroll
  for (int i = 1; i < N; ++i) p ^= par4<LW>(xs, w[i], sel);
  return p;
}
// N words (multiple of 4) from a 16-byte aligned, warp-uniform address
template <int N>
__device__ __forceinline__ void ld_words(const uint32_t* __restrict__ b, uint32_t (&w)[N]) {
  static_assert(N % 4 == 0 || N == 2, "items are 16-byte multiples, or 8 bytes");
  if constexpr (N == 2) {
    const uint2 v = *reinterpret_cast<const uint2*>(b);
    w[0] = v.x; w[1] = v.y;
    return;
  }
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(b + i);
    w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
  }
}

enum SlicedOp { OP_FIRST = 0, OP_LIN = 1, OP_PI = 2, OP_PAIRGEN = 3, OP_PAIRMON = 4 };
enum SlicedRun {
  RUN_LIN = 0, RUN_PI = 3, RUN_LIN2 = 9, RUN_PAIR = 12, RUN_GENERIC = 15,
  // one-word parities (<= 4 rows) in compact items (pack_sliced.py::_emit_runs)
  RUN_LIN_1 = 16, RUN_LIN2_1 = 17, RUN_PI_1 = 18, RUN_PAIR_1 = 21,
  RUN_GENERIC_PI = 22  // generic block stream of pi terms only (masks heavier than 16 rows): FIRST + PI block pairs
};

template <class LW>
struct Planes {
  LW A0, A1, A2, Z;
  LW Bp[5];
};

template <class LW>
__device__ __forceinline__ void lin_op(Planes<LW>& P, uint32_t prm, LW p) {
  add_a3(P.A0, P.A1, P.A2, prm & 7u, p);
  const uint32_t bm = (prm >> 3) & 3u, zm = (prm >> 5) & 3u;
  if (bm) add_cnt5(P.Bp, bm == 1u ? p : ~p);
  if (zm) P.Z |= (zm == 1u ? p : ~p);
}

// The items of a run, IW words each, with the record words of the next item fetched while the current one waits for
// its rows (PF: two item buffers in ping-pong; one shared-memory round trip per item instead of two).  The fetch past
// the last item reads whatever follows the run inside the stage (the chunk ends with the decode tables) and is dropped.
// The narrow layout runs at 72 registers per thread and keeps the plain loop.
template <int IW, bool PF, class F>
__device__ __forceinline__ const uint32_t* run_items(const uint32_t* __restrict__ b, uint32_t count, F&& body) {
  if constexpr (PF) {
    uint32_t wa[IW], wb[IW];
    ld_words<IW>(b, wa);
    uint32_t i = 0;
#pragma unroll 1
    for (;;) {
      ld_words<IW>(b + IW, wb);
      body(wa);
      b += IW;
      if (++i == count) break;
      ld_words<IW>(b + IW, wa);
      body(wb);
      b += IW;
      if (++i == count) break;
    }
  } else {
#pragma unroll 1
    for (uint32_t i = 0; i < count; ++i, b += IW) {
      uint32_t w[IW];
      ld_words<IW>(b, w);
      body(w);
    }
  }
  return b;
}

// LIN run: item = [params, index words]; NW = index words used, item = 4 words (NW <= 3) or 8 words
template <class LW, int NW, bool PF>
__device__ __forceinline__ const uint32_t* lin_run(const uint32_t* __restrict__ b, uint32_t count, uint32_t xs, const uint4& sel, Planes<LW>& P) {
  constexpr int IW = NW <= 3 ? 4 : 8;
  return run_items<IW, PF>(b, count, [&](const uint32_t(&w)[IW]) { lin_op(P, w[0], par_words<LW, NW>(xs, w + 1, sel)); });
}

// PI run: item = [4 index words psi, 4 index words phi], N1 / N2 of them used
template <class LW, int N1, int N2, bool PF>
__device__ __forceinline__ const uint32_t* pi_run(const uint32_t* __restrict__ b, uint32_t count, uint32_t xs, const uint4& sel, Planes<LW>& P) {
  return run_items<8, PF>(
      b, count, [&](const uint32_t(&w)[8]) { P.A2 ^= par_words<LW, N1>(xs, w, sel) & par_words<LW, N2>(xs, w + 4, sel); });
}

// two-parity ops of the phase-pair family
template <class LW>
__device__ __forceinline__ void pair_op(Planes<LW>& P, uint32_t op, uint32_t prm, LW q, LW p, uint32_t nb, LW* __restrict__ plw) {
  if (op == OP_PAIRGEN) {
    const uint32_t r0 = 4u + nb + 2u * (prm & 15u);
    plw[r0 * 32u] = q;
    plw[(r0 + 1u) * 32u] = p;
    return;
  }
  // rare items (a few per graph): rolled loops keep the code small (the kernel is ~100 KB of SASS, instruction fetch counts)
#pragma unroll 1
  for (int v = 0; v < 3; ++v) {
    const LW wv = v == 0 ? q : v == 1 ? p : (q & p);
    add_a3(P.A0, P.A1, P.A2, (prm >> (6 * v)) & 7u, wv);
    const int db = (int)((prm >> (6 * v + 3)) & 7u) - 3;
    const LW w = db > 0 ? wv : ~wv;
#pragma unroll 1
    for (int r = 0; r < (db < 0 ? -db : db); ++r) add_cnt5(P.Bp, w);
  }
  const uint32_t ztt = (prm >> 18) & 15u;
  if (ztt & 1u) P.Z |= ~q & ~p;
  if (ztt & 2u) P.Z |= q & ~p;
  if (ztt & 4u) P.Z |= ~q & p;
  if (ztt & 8u) P.Z |= q & p;
}

// LIN2 run: a += 2 p and nothing else
template <class LW, int NW, bool PF>
__device__ __forceinline__ const uint32_t* lin2_run(const uint32_t* __restrict__ b, uint32_t count, uint32_t xs, const uint4& sel, Planes<LW>& P) {
  constexpr int IW = NW <= 3 ? 4 : 8;
  return run_items<IW, PF>(b, count, [&](const uint32_t(&w)[IW]) {
    const LW p = par_words<LW, NW>(xs, w + 1, sel);
    P.A2 ^= P.A1 & p;
    P.A1 ^= p;
  });
}

// PAIR run: item = [op | params << 3, -, -, -, 4 index words, 4 index words], NW of each used
template <class LW, int NW>
__device__ __forceinline__ const uint32_t* pair_run(const uint32_t* __restrict__ b, uint32_t count, uint32_t xs, const uint4& sel, Planes<LW>& P,
                                                     uint32_t nb, LW* __restrict__ plw) {
  return run_items<12, false>(b, count, [&](const uint32_t(&w)[12]) {
    const LW q = par_words<LW, NW>(xs, w + 4, sel), p = par_words<LW, NW>(xs, w + 8, sel);
    pair_op(P, w[0] & 7u, w[0] >> 3, q, p, nb, plw);
  });
}

// compact items of one-word parities (<= 4 rows)
template <class LW>
__device__ __forceinline__ const uint32_t* lin1_run(const uint32_t* __restrict__ b, uint32_t count, uint32_t xs, const uint4& sel, Planes<LW>& P) {
  return run_items<2, false>(b, count, [&](const uint32_t(&w)[2]) { lin_op(P, w[0], par4<LW>(xs, w[1], sel)); });
}
template <class LW>
__device__ __forceinline__ const uint32_t* lin2_1_run(const uint32_t* __restrict__ b, uint32_t count, uint32_t xs, const uint4& sel, Planes<LW>& P) {
  return run_items<2, false>(b, count, [&](const uint32_t(&w)[2]) {
    const LW p = par4<LW>(xs, w[1], sel);
    P.A2 ^= P.A1 & p;
    P.A1 ^= p;
  });
}
template <class LW, int N2>
__device__ __forceinline__ const uint32_t* pi1_run(const uint32_t* __restrict__ b, uint32_t count, uint32_t xs, const uint4& sel, Planes<LW>& P) {
  return run_items<4, false>(b, count, [&](const uint32_t(&w)[4]) { P.A2 ^= par4<LW>(xs, w[0], sel) & par_words<LW, N2>(xs, w + 1, sel); });
}
template <class LW>
__device__ __forceinline__ const uint32_t* pair1_run(const uint32_t* __restrict__ b, uint32_t count, uint32_t xs, const uint4& sel, Planes<LW>& P,
                                                      uint32_t nb, LW* __restrict__ plw) {
  return run_items<4, false>(b, count, [&](const uint32_t(&w)[4]) {
    const LW q = par4<LW>(xs, w[1], sel), p = par4<LW>(xs, w[2], sel);
    pair_op(P, w[0] & 7u, w[0] >> 3, q, p, nb, plw);
  });
}

// generic run: a stream of parity blocks (pack_sliced.py::_block) with a per-block dispatch
template <class LW>
__device__ __forceinline__ const uint32_t* generic_run(const uint32_t* __restrict__ b, uint32_t words, uint32_t xs, const uint4& sel, Planes<LW>& P,
                                                        uint32_t nb, LW* __restrict__ plw) {
  const uint32_t* __restrict__ end = b + words;
  LW q = LaneWord<LW>::zero();
  while (b < end) {
    const uint2 h = *reinterpret_cast<const uint2*>(b);
    const uint32_t hdr = h.x, n = h.y;
    LW p = LaneWord<LW>::zero();
    for (uint32_t i = 0; i < n; i += 2) {
      const uint2 w = *reinterpret_cast<const uint2*>(b + 2 + i);
      p ^= par4<LW>(xs, w.x, sel) ^ par4<LW>(xs, w.y, sel);
    }
    b += 2 + n;
    const uint32_t op = hdr & 7u, prm = hdr >> 3;
    if (op == OP_FIRST) {
      q = p;
    } else if (op == OP_PI) {
      P.A2 ^= q & p;
    } else if (op == OP_LIN) {
      lin_op(P, prm, p);
    } else {
      pair_op(P, op, prm, q, p, nb, plw);
    }
  }
  return b;
}

// phase 1: the term stream of one graph (typed runs, pack_sliced.py::_emit_runs) for the lanes of the group ->
// plane rows plw[r * 32]: r = 0 "some factor vanished", 1..3 a, 4..4+nb-1 the b counter, then (pa, pb) of every
// in-table general pair.
// The stream is [main | aux] (pack_sliced.py::_split_streams): MAIN_ONLY stops at the end of the main part, the aux part
// (pi runs, about half of the row loads) is then a helper warp's (sliced_phase1_aux).
template <class LW, bool PF, bool MAIN_ONLY>
__device__ __forceinline__ void sliced_phase1(const uint32_t* __restrict__ cbase, uint32_t rec, const LW* __restrict__ xcol,
                                              LW* __restrict__ plw, const uint4& sel) {
  const uint32_t xs = smem_u32(xcol);
  const uint4 h0 = *reinterpret_cast<const uint4*>(cbase + rec);
  const uint32_t nb = (h0.y >> 8) & 0xFFu;
  Planes<LW> P;
  P.A0 = P.A1 = P.A2 = P.Z = LaneWord<LW>::zero();
#pragma unroll
  for (int i = 0; i < 5; ++i) P.Bp[i] = LaneWord<LW>::zero();
  const uint32_t* __restrict__ b = cbase + rec + kSlicedHeaderWords;
  const uint32_t* __restrict__ end = b + (MAIN_ONLY ? (h0.w >> 16) : (h0.x & 0xFFFFu));
  while (b < end) {
    const uint32_t rh = *b;
    const uint32_t kind = rh & 0xFFFFu, count = rh >> 16;
    b += 4;
    switch (kind) {
      case RUN_LIN + 0: b = lin_run<LW, 2, PF>(b, count, xs, sel, P); break;
      case RUN_LIN + 1: b = lin_run<LW, 3, PF>(b, count, xs, sel, P); break;
      case RUN_LIN + 2: b = lin_run<LW, 4, PF>(b, count, xs, sel, P); break;
      case RUN_PI + 0: b = pi_run<LW, 2, 2, PF>(b, count, xs, sel, P); break;
      case RUN_PI + 1: b = pi_run<LW, 2, 3, PF>(b, count, xs, sel, P); break;
      case RUN_PI + 2: b = pi_run<LW, 2, 4, PF>(b, count, xs, sel, P); break;
      case RUN_PI + 3: b = pi_run<LW, 3, 3, PF>(b, count, xs, sel, P); break;
      case RUN_PI + 4: b = pi_run<LW, 3, 4, PF>(b, count, xs, sel, P); break;
      case RUN_PI + 5: b = pi_run<LW, 4, 4, PF>(b, count, xs, sel, P); break;
      case RUN_LIN2 + 0: b = lin2_run<LW, 2, PF>(b, count, xs, sel, P); break;
      case RUN_LIN2 + 1: b = lin2_run<LW, 3, PF>(b, count, xs, sel, P); break;
      case RUN_LIN2 + 2: b = lin2_run<LW, 4, PF>(b, count, xs, sel, P); break;
      case RUN_PAIR + 0: b = pair_run<LW, 2>(b, count, xs, sel, P, nb, plw); break;
      case RUN_PAIR + 1: b = pair_run<LW, 3>(b, count, xs, sel, P, nb, plw); break;
      case RUN_PAIR + 2: b = pair_run<LW, 4>(b, count, xs, sel, P, nb, plw); break;
      case RUN_LIN_1: b = lin1_run<LW>(b, count, xs, sel, P); break;
      case RUN_LIN2_1: b = lin2_1_run<LW>(b, count, xs, sel, P); break;
      case RUN_PI_1 + 0: b = pi1_run<LW, 1>(b, count, xs, sel, P); break;
      case RUN_PI_1 + 1: b = pi1_run<LW, 2>(b, count, xs, sel, P); break;
      case RUN_PI_1 + 2: b = pi1_run<LW, 3>(b, count, xs, sel, P); break;
      case RUN_PAIR_1: b = pair1_run<LW>(b, count, xs, sel, P, nb, plw); break;
      default: b = generic_run<LW>(b, count, xs, sel, P, nb, plw); break;
    }
  }
  plw[0] = P.Z;
  plw[32] = P.A0;
  plw[64] = P.A1;
  plw[96] = P.A2;
#pragma unroll
  for (int i = 0; i < 5; ++i)
    if ((uint32_t)i < nb) plw[(4 + i) * 32] = P.Bp[i];
}

// phase 1 of a helper warp: the aux part of the stream (pi runs only: A2 ^= q & p, a pure XOR into the top plane of
// `a`) -> plane row kMinPlaneRows of the graph's slot, XORed into A2 by phase 2
template <class LW>
__device__ __forceinline__ void sliced_phase1_aux(const uint32_t* __restrict__ cbase, uint32_t rec, const LW* __restrict__ xcol,
                                                  LW* __restrict__ plw, const uint4& sel) {
  const uint32_t xs = smem_u32(xcol);
  const uint4 h0 = *reinterpret_cast<const uint4*>(cbase + rec);
  Planes<LW> P;
  P.A2 = LaneWord<LW>::zero();
  const uint32_t* __restrict__ b = cbase + rec + kSlicedHeaderWords + (h0.w >> 16);
  const uint32_t* __restrict__ end = cbase + rec + kSlicedHeaderWords + (h0.x & 0xFFFFu);
  while (b < end) {
    const uint32_t rh = *b;
    const uint32_t kind = rh & 0xFFFFu, count = rh >> 16;
    b += 4;
    switch (kind) {
      case RUN_PI + 0: b = pi_run<LW, 2, 2, false>(b, count, xs, sel, P); break;
      case RUN_PI + 1: b = pi_run<LW, 2, 3, false>(b, count, xs, sel, P); break;
      case RUN_PI + 2: b = pi_run<LW, 2, 4, false>(b, count, xs, sel, P); break;
      case RUN_PI + 3: b = pi_run<LW, 3, 3, false>(b, count, xs, sel, P); break;
      case RUN_PI + 4: b = pi_run<LW, 3, 4, false>(b, count, xs, sel, P); break;
      case RUN_PI + 5: b = pi_run<LW, 4, 4, false>(b, count, xs, sel, P); break;
      case RUN_PI_1 + 0: b = pi1_run<LW, 1>(b, count, xs, sel, P); break;
      case RUN_PI_1 + 1: b = pi1_run<LW, 2>(b, count, xs, sel, P); break;
      case RUN_PI_1 + 2: b = pi1_run<LW, 3>(b, count, xs, sel, P); break;
      case RUN_GENERIC_PI: {  // FIRST block (q) + PI block (A2 ^= q & p), masks of any weight
        const uint32_t* __restrict__ gend = b + count;
        LW q = LaneWord<LW>::zero();
        while (b < gend) {
          const uint2 h = *reinterpret_cast<const uint2*>(b);
          LW p = LaneWord<LW>::zero();
          for (uint32_t i = 0; i < h.y; i += 2) {
            const uint2 w = *reinterpret_cast<const uint2*>(b + 2 + i);
            p ^= par4<LW>(xs, w.x, sel) ^ par4<LW>(xs, w.y, sel);
          }
          b += 2 + h.y;
          if ((h.x & 7u) == OP_PI) P.A2 ^= q & p;
          else q = p;  // OP_FIRST (and the zero padding at the end of the run)
        }
        break;
      }
      default: b = end; break;  // the packer puts nothing else here
    }
  }
  plw[kMinPlaneRows * 32] = P.A2;
}

template <bool HAS_EXACT>
struct SlicedAcc { typedef float2 type; };
template <>
struct SlicedAcc<true> { typedef uint4 type; };

// 8 x 8 bit transpose: in, byte k of (lo | hi << 32) = plane k (bit s = shot s); out, byte s = index of shot s (bit k)
__device__ __forceinline__ void transpose8x8(uint32_t& lo, uint32_t& hi) {
  uint32_t t;
  t = (lo ^ (lo >> 7)) & 0x00AA00AAu; lo ^= t ^ (t << 7);
  t = (hi ^ (hi >> 7)) & 0x00AA00AAu; hi ^= t ^ (t << 7);
  t = (lo ^ (lo >> 14)) & 0x0000CCCCu; lo ^= t ^ (t << 14);
  t = (hi ^ (hi >> 14)) & 0x0000CCCCu; hi ^= t ^ (t << 14);
  t = (lo ^ (hi << 4)) & 0xF0F0F0F0u; lo ^= t; hi ^= t >> 4;
}

// phase 2: this warp's SH shots of every slab (warp w of the group: shots w * SH ...): the plane bits of a shot form
// the index of its decode-table entry.  The first eight index planes are gathered with byte permutes and one 8 x 8
// bit transpose, and a dp4a per shot scales the index byte into the entry's address; further planes are rare.
// plj: this lane's word of plane row 0 (rows are 32 lane words apart); a wide lane word is read once for both slabs.
__device__ __forceinline__ uint32_t lw_half(uint32_t v, int) { return v; }
__device__ __forceinline__ uint32_t lw_half(const U2& v, int h) { return h ? v.y : v.x; }

template <int SH, bool HAS_EXACT, class LW, bool HELP>
__device__ __forceinline__ void sliced_phase2(const uint32_t* __restrict__ cbase, uint32_t rec, const LW* __restrict__ plj, int w,
                                              bool approx, const uint4& sel_e, typename SlicedAcc<HAS_EXACT>::type* __restrict__ acc_all,
                                              const int4* __restrict__ pair_tab) {
  static_assert(SH == 8 || SH == 4, "a warp handles a byte or a nibble of every plane word");
  constexpr int NH = LaneWord<LW>::kWords;
  const uint4 h0 = *reinterpret_cast<const uint4*>(cbase + rec);
  const int n_idx = (int)(h0.y & 0xFFu);
  const uint32_t tbl = smem_u32(cbase + h0.z);
  constexpr uint32_t kField = (1u << SH) - 1u;
  // entries are 8 bytes (float2) in approximate levels and 16 bytes (four coefficients) in exact ones
  const uint32_t kEntryShift = (HAS_EXACT && !approx) ? 4u : 3u;
  const int sh0 = w * SH;
  const uint32_t bsel = (uint32_t)(SH == 8 ? w : (w >> 1));  // byte of the plane words that holds this warp's shots
  const uint32_t ps = bsel | ((4u + bsel) << 4);
  const uint32_t zero_entry = smem_u32(cbase + rec + 4);  // reserved header words: a shot whose value vanished adds 0
  LW pbw[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) pbw[k] = k < n_idx ? plj[(1 + k) * 32] : LaneWord<LW>::zero();
  if constexpr (HELP) pbw[2] ^= plj[kMinPlaneRows * 32];  // the helper warp's share of A2 (pi runs of the aux part)
  const LW zw = plj[0];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    typename SlicedAcc<HAS_EXACT>::type* __restrict__ acc = acc_all + h * SH;
    uint32_t pb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pb[k] = lw_half(pbw[k], h);
    uint32_t lo = __byte_perm(__byte_perm(pb[0], pb[1], ps), __byte_perm(pb[2], pb[3], ps), 0x5410);
    uint32_t hi = __byte_perm(__byte_perm(pb[4], pb[5], ps), __byte_perm(pb[6], pb[7], ps), 0x5410);
    transpose8x8(lo, hi);
    uint32_t base[SH];
#pragma unroll
    for (int s = 0; s < SH; ++s) base[s] = tbl;
    for (int k = 8; k < n_idx; ++k) {
      const uint32_t pk = ((lw_half(plj[(1 + k) * 32], h) >> sh0) & kField) << (k + kEntryShift);
      const uint32_t m = 1u << (k + kEntryShift);
#pragma unroll
      for (int s = 0; s < SH; ++s) base[s] += (pk >> s) & m;
    }
    const uint32_t zb = lw_half(zw, h) >> sh0;
    const uint32_t r0 = SH == 8 ? lo : ((w & 1) ? hi : lo), r1 = hi;
    // exact levels: general phase pairs applied as ring factors (two-stage decode, pack_sliced.py): pair j's parities
    // sit in planes 1 + n_idx + 2 j (alpha side) and + 1 (beta side); its control byte alpha | beta << 3 in the record's
    // last four words.  qab[j] = this warp's SH alpha bits | SH beta bits << 8.
    uint32_t n_mul = 0, ctlw = 0, qab[4] = {0u, 0u, 0u, 0u};
    if constexpr (HAS_EXACT) {
      if (!approx) {
        n_mul = (h0.y >> 16) & 0xFFu;
        if (n_mul) {
          ctlw = cbase[rec + (h0.w & 0xFFFFu) - 4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((uint32_t)j < n_mul)
              qab[j] = ((lw_half(plj[(1 + n_idx + 2 * j) * 32], h) >> sh0) & kField) |
                       (((lw_half(plj[(2 + n_idx + 2 * j) * 32], h) >> sh0) & kField) << 8);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < SH; ++s) {
      const uint32_t r = s < 4 ? r0 : r1;
      uint32_t sl = (s & 3) == 0 ? sel_e.x : (s & 3) == 1 ? sel_e.y : (s & 3) == 2 ? sel_e.z : sel_e.w;
      if (HAS_EXACT && !approx) sl <<= 1;
      uint32_t ea = __dp4a(r, sl, base[s]);
      ea = ((zb >> s) & 1u) ? zero_entry : ea;
      if (approx) {
        const float2 e = *reinterpret_cast<const float2*>(__cvta_shared_to_generic(ea));
        if constexpr (HAS_EXACT) {
          acc[s].x = __float_as_uint(__fadd_rn(__uint_as_float(acc[s].x), e.x));
          acc[s].y = __float_as_uint(__fadd_rn(__uint_as_float(acc[s].y), e.y));
        } else {
          acc[s].x = __fadd_rn(acc[s].x, e.x);
          acc[s].y = __fadd_rn(acc[s].y, e.y);
        }
      } else {
        if constexpr (HAS_EXACT) {
          uint4 e = *reinterpret_cast<const uint4*>(__cvta_shared_to_generic(ea));
          if (n_mul) {  // warp-uniform
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if ((uint32_t)j < n_mul) {
                const uint32_t ix = ((ctlw >> (8 * j)) & 63u) ^ (((qab[j] >> s) & 1u) << 2) ^ (((qab[j] >> (8 + s)) & 1u) << 5);
                const ZW pr = zw_mul(ZW{e.x, e.y, e.z, e.w}, zw_from(pair_tab[ix]));
                e = make_uint4(pr.c0, pr.c1, pr.c2, pr.c3);
              }
            }
          }
          acc[s].x += e.x; acc[s].y += e.y; acc[s].z += e.z; acc[s].w += e.w;
        }
      }
    }
  }
}

__host__ __device__ constexpr int sliced_max_groups(int split) { return split == 4 ? 7 : 3; }
constexpr int kWideMaxGroups = 4;  // wide layout: 4 warps per group, up to 8 units of 1024 shots per CTA
// HELP: one group of 8 main + 8 helper warps (thin launches)
__host__ __device__ constexpr int sliced_max_threads(int split, bool wide = false, bool help = false) {
  return help ? 2 * split * 32 : (wide ? kWideMaxGroups : sliced_max_groups(split)) * split * 32;
}

__device__ __forceinline__ void group_sync(int grp, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(nthreads) : "memory");
}
// stages of the ring: one mbarrier ("full") and one named barrier (8 + stage, "empty") each; the groups use named barriers 1..7
constexpr int kSlicedMaxStages = 8;

// dynamic shared memory (32-bit words): [0,64) mbarriers | xt [groups][rows][32 lanes] | planes
// [groups][2][SPLIT][plane_rows][32 lanes] | stage ring; a lane is one word (narrow) or two (wide)
// HELP (thin launches: one group per SM, bound by the latency of one warp's walk through a graph): the group gets SPLIT
// helper warps; helper w runs the aux part of graph wave + w (pi runs, about half of the row loads) while main warp w
// runs the main part, and phase 2 XORs the helper's plane into A2.  Helpers skip phase 2, so they are already in the
// next wave's phase 1 while the main warps decode.
template <int SPLIT, bool HAS_EXACT, bool HAS_ROWS, bool WIDE, bool HELP>
__global__ void __launch_bounds__(sliced_max_threads(SPLIT, WIDE, HELP), 1) sample_sliced_kernel(const SParams prm) {
  static_assert(!WIDE || (SPLIT == 4 && !HAS_EXACT), "the wide layout is built for the 4-way split of approximate programs");
  static_assert(!HELP || (SPLIT == 8 && !HAS_EXACT && !WIDE), "helper warps serve the 8-way split of approximate programs");
  constexpr int GW = HELP ? 2 * SPLIT : SPLIT;  // warps per group
  constexpr int SH = 32 / SPLIT;
  constexpr int NH = WIDE ? 2 : 1;  // slabs per lane
  // Fetching the record of the next item while the current one waits for its rows (run_items PF) was measured twice and
  // buys nothing: wide layout at 128 registers 0.727 -> 0.738 ms, thin 8-way launches 0.311 -> 0.311 ms per memoised step.
  constexpr bool kPrefetchRecords = false;
  typedef typename std::conditional<WIDE, U2, uint32_t>::type LW;
  typedef LaneWord<LW> L;
  // a *unit* is 32 slabs (1024 shots): a narrow group, or half the lanes of a wide group
  int n_slabs = prm.n_slabs, n_units = prm.n_groups, rounds = prm.rounds;
  long long n_live = prm.B;  // slots of the launch (batch rows, or entries of the row list)
  if constexpr (HAS_ROWS) {
    const long long nr = (long long)*prm.n_rows;
    n_live = nr;
    n_slabs = (int)((nr + 31) / 32);
    n_units = (n_slabs + 31) / 32;
    if ((int)blockIdx.x >= n_units) return;  // this CTA owns no unit in any round
    const int upc = (n_units + (int)gridDim.x - 1) / (int)gridDim.x;
    rounds = (upc + prm.ng - 1) / prm.ng;
  }
  typedef typename SlicedAcc<HAS_EXACT>::type Acc;
  extern __shared__ __align__(128) uint32_t smem[];
  const uint32_t* __restrict__ blob = prm.blob;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int grp = wid / GW, w = wid % SPLIT;
  const bool helper = HELP && (wid % GW) >= SPLIT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // one graph slot of a plane buffer (lane words); only exact levels multiply pairs, so all-approximate programs keep the constant
  const int plane_words = HAS_EXACT ? prm.plane_rows * 32 : (kMinPlaneRows + (HELP ? 1 : 0)) * 32;
  const uint4 sel = prm.sel;
  uint32_t* sdata = smem + prm.smem_data_off;

  const int n_comp = (int)blob[H_N_COMP], n_chunks = (int)blob[H_N_CHUNKS];
  const uint32_t* tabs = blob + blob[H_OFF_COMP];  // component | level | chunk tables, contiguous in the blob
  if (prm.smem_tab_off >= 0) {
    uint32_t* st = smem + prm.smem_tab_off;
    for (int i = tid; i < prm.tab_words; i += (int)blockDim.x) st[i] = tabs[i];
    tabs = st;  // visible after the __syncthreads below
  }
  const uint32_t* __restrict__ comp_tab = tabs;
  const uint32_t* __restrict__ level_tab = tabs + (blob[H_OFF_LEVEL] - blob[H_OFF_COMP]);
  const uint32_t* __restrict__ chunk_tab = tabs + (blob[H_OFF_CHUNK] - blob[H_OFF_COMP]);
  const uint32_t* __restrict__ gdata = blob + blob[H_OFF_DATA];
  const int one_row = (int)blob[H_ONE_ROW];

  // 1 + w^a + w^b - w^(a+b), index a | b << 3 (terms.py:179-181): factors of the multiplied general pairs (exact levels)
  __shared__ int4 s_pair[HAS_EXACT ? 64 : 1];
  if constexpr (HAS_EXACT) {
    for (int i = tid; i < 64; i += (int)blockDim.x) {
      const int a = i & 7, b = i >> 3;
      const int4 ua = unit_phase(a), ub = unit_phase(b), uc = unit_phase(a + b);
      s_pair[i] = make_int4(1 + ua.x + ub.x - uc.x, ua.y + ub.y - uc.y, ua.z + ub.z - uc.z, ua.w + ub.w - uc.w);
    }
  }
  // Ring flow control without a CTA-wide barrier (prm.lockstep == 0): a warp that is through with a stage *arrives* on the
  // stage's named barrier and carries on; warp 0 *waits* on it and refills the stage (split barrier: bar.arrive /
  // bar.sync).  Groups therefore drift apart by up to n_stages chunks, so that one group's phase 2 and level ends
  // (issue-bound) overlap another group's phase 1 (shared-memory bound) instead of all groups moving in lockstep.
  if (tid == 0) {
    for (int i = 0; i < prm.n_stages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();

  // chunk fills are numbered q = 0 .. rounds * n_chunks - 1 (fits 32 bits: the host plans at most 2^20 rounds)
  const uint32_t total_q = (uint32_t)rounds * (uint32_t)n_chunks;
  auto issue = [&](uint32_t q) {
    const uint32_t ch = q % (uint32_t)n_chunks;
    const uint32_t* row = chunk_tab + ch * kChunkWords;
    const uint32_t stage = q % (uint32_t)prm.n_stages;
    const uint32_t bytes = row[K_WORDS] * 4u;
    mbar_expect_tx(&bars[stage], bytes);
    tma_bulk_g2s(sdata + (size_t)stage * prm.stage_words, gdata + row[K_OFF], bytes, &bars[stage]);
  };
  if (tid == 0)
    for (uint32_t q = 0; q < (uint32_t)prm.n_stages && q < total_q; ++q) issue(q);

  uint32_t q = 0, q_stage = 0, q_phase = 0;  // q_stage = q % n_stages, q_phase = (q / n_stages) & 1, kept incrementally
  uint32_t pbuf = 0;  // plane buffer of the current wave (double-buffered: one group barrier per wave)
  for (int round = 0; round < rounds; ++round) {
    // units of this group in this round, the lane's slabs and its column of the group's matrix.  Lanes 16..31 of a
    // half-filled wide group sit out (owner = false): 64-bit shared-memory accesses are served per half-warp, so an
    // idle half costs no wavefront, and the group's traffic stays proportional to its shots.
    int slab0, unit0 = 0;
    const int col = lane;
    bool gactive, owner = true;
    if constexpr (WIDE) {
      const int cu = 2 * grp;  // first unit of the group inside the CTA
      const int u0 = (round * prm.ng + cu) * (int)gridDim.x + (int)blockIdx.x;
      const int u1 = (round * prm.ng + cu + 1) * (int)gridDim.x + (int)blockIdx.x;
      gactive = cu < prm.ng && u0 < n_units;
      const bool have1 = cu + 1 < prm.ng && u1 < n_units;
      const bool upper = lane >= 16;
      if (upper && !have1) owner = false;
      slab0 = ((upper && have1) ? u1 : u0) * 32 + 2 * (lane & 15);
    } else {
      const int u0 = (round * prm.ng + grp) * (int)gridDim.x + (int)blockIdx.x;
      gactive = u0 < n_units;  // uniform over the group's warps
      slab0 = u0 * 32 + lane;
      unit0 = u0;
    }
    if (helper) owner = false;  // helpers take no part in the main warps' work (fills, phase 2, level ends, row assembly)
    bool active[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) active[h] = gactive && slab0 + h < n_slabs;
    // word offsets of this thread's columns; made opaque so that the compiler keeps them in registers instead of
    // re-deriving them from threadIdx inside the block loop (it did, ten instructions per block)
    uint32_t xoff = grp * prm.rows * 32 + col;
    uint32_t ploff = grp * (2 * SPLIT * plane_words) + col;  // two plane buffers per group
    asm volatile("" : "+r"(xoff), "+r"(ploff));
    LW* xcol = reinterpret_cast<LW*>(smem + prm.smem_xt_off) + xoff;
    LW* plg = reinterpret_cast<LW*>(smem + prm.smem_pl_off) + ploff;

    int xt_row0 = 0, draw0 = 0;
    for (int ci = 0; ci < n_comp; ++ci) {
      const uint32_t* __restrict__ comp = comp_tab + ci * kCompWords;
      const int F = (int)comp[C_F], n_c = (int)comp[C_NC];
      if (gactive) {
        group_sync(grp, GW * 32);  // the previous component's readers are done with the columns
        bool fused_in = false;
        if constexpr (!WIDE) fused_in = prm.f_rows != nullptr;
        if (fused_in) {
          if (!helper) {
           if constexpr (!WIDE) {
            // K0t inside the group: warp w turns the f rows of slabs w * SPW .. of the unit into matrix columns (lane =
            // shot; one ballot per selected f bit gives the row word of the slab).  The f words of all the warp's slabs
            // are fetched first (one DRAM round trip), the selection table travels lane to lane by shuffles (no dependent
            // global loads in the bit loop).  Stores of one slab hit one bank: ~50 wavefronts per slab and component,
            // against ~4 500 row loads per slab in phase 1.
            for (int i = F + w; i < prm.rows; i += SPLIT) xcol[i * 32] = i == one_row ? 0xFFFFFFFFu : 0u;
            const uint32_t* __restrict__ fsel = blob + blob[H_OFF_FSEL] + comp[C_FSEL_OFF];
            const int wf = (int)blob[H_WF64];  // <= 4 (host: wider rows keep the separate kernels)
            uint32_t* xg = xcol - lane;
            constexpr int SPW = 32 / SPLIT;
            auto row_ptr = [&](int j) -> const uint64_t* {  // f row of this lane's shot in slab j of the warp, or null
              const long long slot = ((long long)unit0 * 32 + w * SPW + j) * 32 + lane;
              if (slot >= n_live) return nullptr;
              const long long r = HAS_ROWS ? (long long)prm.row_list[slot] : slot;
              return prm.f_rows + r * wf;
            };
            // word 0 of the slabs' rows is fetched two slabs ahead (DRAM latency); the slab loop stays rolled: this
            // runs once per component and must not bloat the kernel's code
            auto word0 = [&](int j) -> uint64_t {
              const uint64_t* fr = j < SPW ? row_ptr(j) : nullptr;
              return fr ? fr[0] : 0ull;
            };
            uint64_t f_cur = word0(0), f_n1 = word0(1);
#pragma unroll 1
            for (int j = 0; j < SPW; ++j) {
              const uint64_t f_n2 = word0(j + 2);
              uint64_t w1 = 0, w2 = 0, w3 = 0;  // further words of wide f rows: fetched where needed (L1 / L2)
              if (wf > 1) {
                const uint64_t* fr = row_ptr(j);
                if (fr) {
                  w1 = fr[1];
                  if (wf > 2) w2 = fr[2];
                  if (wf > 3) w3 = fr[3];
                }
              }
              const uint32_t f_lo = (uint32_t)f_cur, f_hi = (uint32_t)(f_cur >> 32);
              uint32_t* xs_slab = xg + w * SPW + j;  // column of this slab; lane 0 stores the row words as they come
#pragma unroll 1
              for (int i0 = 0; i0 < F; i0 += 32) {
                const int lim = min(32, F - i0);
                const uint32_t fi_lane = lane < lim ? fsel[i0 + lane] : 0u;
                if (wf == 1) {  // the common case (num_f <= 64): 32-bit selects and shifts only
#pragma unroll 8
                  for (int jj = 0; jj < lim; ++jj) {
                    const uint32_t fi = __shfl_sync(0xFFFFFFFFu, fi_lane, jj);
                    const uint32_t half = (fi & 32u) ? f_hi : f_lo;
                    const uint32_t word = __ballot_sync(0xFFFFFFFFu, ((half >> (fi & 31u)) & 1u) != 0u);
                    if (lane == 0) xs_slab[(i0 + jj) * 32] = word;
                  }
                } else {
#pragma unroll 4
                  for (int jj = 0; jj < lim; ++jj) {
                    const uint32_t fi = __shfl_sync(0xFFFFFFFFu, fi_lane, jj);
                    const uint32_t wsel = fi >> 6;
                    const uint64_t wv = wsel == 0 ? f_cur : wsel == 1 ? w1 : wsel == 2 ? w2 : w3;
                    const uint32_t word = __ballot_sync(0xFFFFFFFFu, ((uint32_t)(wv >> (fi & 63u)) & 1u) != 0u);
                    if (lane == 0) xs_slab[(i0 + jj) * 32] = word;
                  }
                }
              }
              f_cur = f_n1;
              f_n1 = f_n2;
            }
           }
          }
        } else {
          for (int i = w; i < prm.rows; i += SPLIT) {
            LW v = L::zero();
            if (i < F) {
              const uint32_t* src = prm.xt + (size_t)(xt_row0 + i) * prm.slab_cap + slab0;
              if constexpr (WIDE) {
                v.x = active[0] ? src[0] : 0u;
                v.y = active[1] ? src[1] : 0u;
              } else {
                v = active[0] ? src[0] : 0u;
              }
            } else if (i == one_row) {
              v = L::ones();
            }
            if (owner) xcol[i * 32] = v;
          }
        }
      }

      for (int k = 0; k <= n_c; ++k) {
        const uint32_t* __restrict__ lvl = level_tab + (comp[C_FIRST_LEVEL] + k) * kLevelWords;
        const bool approx = (lvl[L_FLAGS] & 1u) != 0u;
        if (gactive) {
          if (k > 0 && w == 0 && owner) xcol[(F + k - 1) * 32] = L::ones();  // trying bit 1 for every shot
          group_sync(grp, GW * 32);
        }
        Acc acc[NH * SH];
#pragma unroll
        for (int s = 0; s < NH * SH; ++s) {
          if constexpr (HAS_EXACT) acc[s] = make_uint4(0u, 0u, 0u, 0u);
          else acc[s] = make_float2(0.0f, 0.0f);
        }
        const int first_chunk = (int)lvl[L_FIRST_CHUNK], nck = (int)lvl[L_N_CHUNKS];
        for (int c = 0; c < nck; ++c) {
          const uint32_t* row = chunk_tab + (first_chunk + c) * kChunkWords;
          const uint32_t stage = q_stage;
          mbar_wait(&bars[stage], q_phase);
          const uint32_t* __restrict__ cbase = sdata + (size_t)stage * prm.stage_words;
          const int n_g = (int)row[K_GRAPHS];
          if (gactive) {
            for (int w0 = 0; w0 < n_g; w0 += SPLIT) {
              // A warp may write the other buffer for the next wave as soon as it is through with this one: everybody
              // passed this wave's barrier, hence finished reading that buffer in the wave before.
              LW* pw = plg + pbuf * (SPLIT * plane_words);
              if (w0 + w < n_g) {
                if (owner) sliced_phase1<LW, kPrefetchRecords, HELP>(cbase, cbase[w0 + w], xcol, pw + w * plane_words, sel);
                if constexpr (HELP) {
                  if (helper) sliced_phase1_aux<LW>(cbase, cbase[w0 + w], xcol, pw + w * plane_words, sel);
                }
              }
              group_sync(grp, GW * 32);
              const int nj = min(SPLIT, n_g - w0);
              if (owner)
                for (int j = 0; j < nj; ++j)
                  sliced_phase2<SH, HAS_EXACT, LW, HELP>(cbase, cbase[w0 + j], pw + j * plane_words, w, approx, prm.sel_e, acc, s_pair);
              pbuf ^= 1u;
            }
          }
          if (prm.lockstep) {  // CTA-wide barrier per chunk: all groups walk the same code at the same time
            __syncthreads();
            if (tid == 0 && q + prm.n_stages < total_q) issue(q + prm.n_stages);
          } else {
            // split barrier per stage: every warp arrives when it is through with the stage, warp 0 waits for all of
            // them and refills.  Named barriers 8.. (the groups use 1..7); the other warps do not wait.
            if (wid == 0) {
              asm volatile("bar.sync %0, %1;" ::"r"(8 + (int)stage), "r"((int)blockDim.x) : "memory");
              if (tid == 0 && q + prm.n_stages < total_q) {
                fence_proxy_async();
                issue(q + prm.n_stages);
              }
            } else {
              asm volatile("bar.arrive %0, %1;" ::"r"(8 + (int)stage), "r"((int)blockDim.x) : "memory");
            }
          }
          ++q;
          if (++q_stage == (uint32_t)prm.n_stages) {
            q_stage = 0;
            q_phase ^= 1u;
          }
        }
        // finish the level for this warp's shots: |amp|, draw, chain rule
        if (gactive) {
          const int p_lo = (int)lvl[L_P_LO];
          const bool empty = lvl[L_G] == 0u;
          uint32_t k0 = 0, k1 = 0;
          if (k > 0) {
            k0 = prm.subkeys[2 * (draw0 + k - 1)];
            k1 = prm.subkeys[2 * (draw0 + k - 1) + 1];
          }
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            if (!owner) break;
            const int slab = slab0 + h;
            const unsigned long long shot0 =
                (unsigned long long)prm.shot_offset + (unsigned long long)slab * 32ull + (unsigned long long)(w * SH);
            // chain-rule state of this warp's shots lives in global memory between levels (L2-resident, 32 B per lane)
            float4* pvp = reinterpret_cast<float4*>(prm.pv + (size_t)slab * 32 + w * SH);
            float pvv[SH];
#pragma unroll
            for (int s = 0; s < SH; s += 4) {
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (k > 0 && active[h]) t = pvp[s >> 2];
              pvv[s] = t.x; pvv[s + 1] = t.y; pvv[s + 2] = t.z; pvv[s + 3] = t.w;
            }
            uint32_t rid[SH];  // HAS_ROWS: batch row (= RNG counter) of each shot of this lane's slab
            if constexpr (HAS_ROWS) {
              const uint4* rp = reinterpret_cast<const uint4*>(prm.row_list + (size_t)slab * 32 + w * SH);
#pragma unroll
              for (int s = 0; s < SH; s += 4) {
                uint4 t = make_uint4(0u, 0u, 0u, 0u);
                if (k > 0 && active[h]) t = rp[s >> 2];
                rid[s] = t.x; rid[s + 1] = t.y; rid[s + 2] = t.z; rid[s + 3] = t.w;
              }
            }
            uint32_t bits = 0;
#pragma unroll
            for (int s = 0; s < SH; ++s) {
              const Acc a = acc[h * SH + s];
              float re = 0.0f, im = 0.0f;
              if (approx) {
                if constexpr (HAS_EXACT) { re = __uint_as_float(a.x); im = __uint_as_float(a.y); }
                else { re = a.x; im = a.y; }
              } else {
                if constexpr (HAS_EXACT) {
                  ZW cz = ZW{a.x, a.y, a.z, a.w};
                  int p = p_lo;
                  zw_fixpoint(cz, p);
                  zw_to_complex(cz, p, re, im);
                }
              }
              if (empty) { re = 0.0f; im = 0.0f; }
              const float p1 = complex_abs(re, im);
              if (k == 0) {
                pvv[s] = p1;
              } else {
                const float pv = pvv[s];
                const unsigned long long ctr = HAS_ROWS ? (unsigned long long)prm.shot_offset + (unsigned long long)rid[s]
                                                        : shot0 + (unsigned long long)s;
                const float u = uniform_f32(k0, k1, ctr);
                const bool bit = u < __fdiv_rn(p1, pv);
                pvv[s] = bit ? p1 : __fsub_rn(pv, p1);
                bits |= (bit ? 1u : 0u) << s;
              }
            }
            if (active[h] && k < n_c) {
#pragma unroll
              for (int s = 0; s < SH; s += 4) pvp[s >> 2] = make_float4(pvv[s], pvv[s + 1], pvv[s + 2], pvv[s + 3]);
            }
            if (k > 0) {
              constexpr uint32_t kField = SH == 32 ? 0xFFFFFFFFu : ((1u << SH) - 1u);
              atomicAnd(reinterpret_cast<uint32_t*>(&xcol[(F + k - 1) * 32]) + h, (bits << (w * SH)) | ~(kField << (w * SH)));
            }
          }
          if (k > 0) {
            group_sync(grp, GW * 32);
            if (w == 0 && owner) {
              const uint32_t* drawn = reinterpret_cast<const uint32_t*>(&xcol[(F + k - 1) * 32]);
#pragma unroll
              for (int h = 0; h < NH; ++h)
                if (active[h]) prm.ot[(size_t)(draw0 + k - 1) * prm.slab_cap + slab0 + h] = drawn[h];
            }
          }
        }
      }
      xt_row0 += F;
      draw0 += n_c;
    }
    if constexpr (!WIDE) {
      if (prm.out_rows != nullptr && gactive) {
        // K2a inside the group: packed output rows of the unit's shots (direct bits of f + the drawn bits, which the
        // group has just written to ot).  Warp w takes slabs w * SPW ..; lane = shot.  Table entries (direct_tab, dest)
        // and the slab's drawn words are loaded one per lane and passed around by shuffles.
        group_sync(grp, GW * 32);
        const int wf = (int)blob[H_WF64], wo = (int)blob[H_WOUT64];
        const int n_direct = (int)blob[H_N_DIRECT], n_draws = (int)blob[H_N_DRAWS];
        const uint32_t* __restrict__ direct_tab = blob + blob[H_OFF_DIRECT];
        const uint32_t* __restrict__ dest = blob + blob[H_OFF_DEST];
        constexpr int SPW = 32 / SPLIT;
        for (int j = 0; j < (helper ? 0 : SPW); ++j) {
          const long long gslab = (long long)unit0 * 32 + w * SPW + j;
          const long long slot = gslab * 32 + lane;
          const bool valid = slot < n_live;
          long long r = 0;
          if (valid) r = HAS_ROWS ? (long long)prm.row_list[slot] : slot;
          const uint64_t* __restrict__ frow = prm.f_rows + r * wf;
          uint64_t v0 = 0, v1 = 0;  // output words 0 and 1 (host: wider rows keep the separate kernel)
          if (wf == 1 && wo == 1) {  // the common case: one f word, one output word, 32-bit halves throughout
            const uint64_t f0 = valid ? frow[0] : 0ull;
            const uint32_t f_lo = (uint32_t)f0, f_hi = (uint32_t)(f0 >> 32);
            uint32_t v_lo = 0, v_hi = 0;
            for (int jd0 = 0; jd0 < n_direct; jd0 += 32) {
              const int lim = min(32, n_direct - jd0);
              const uint32_t fi_lane = lane < lim ? direct_tab[2 * (jd0 + lane)] : 0u;
              const uint32_t dd_lane = lane < lim ? direct_tab[2 * (jd0 + lane) + 1] : 0u;
#pragma unroll 4
              for (int jd = 0; jd < lim; ++jd) {
                const uint32_t fi = __shfl_sync(0xFFFFFFFFu, fi_lane, jd), dd = __shfl_sync(0xFFFFFFFFu, dd_lane, jd);
                const uint32_t bit = ((((fi & 32u) ? f_hi : f_lo) >> (fi & 31u)) & 1u) ^ (dd >> 31);
                if (dd & 32u) v_hi |= bit << (dd & 31u);  // warp-uniform
                else v_lo |= bit << (dd & 31u);
              }
            }
            for (int jd0 = 0; jd0 < n_draws; jd0 += 32) {
              const int lim = min(32, n_draws - jd0);
              const uint32_t d_lane = lane < lim ? dest[jd0 + lane] : 0u;
              const uint32_t o_lane = (lane < lim && gslab < n_slabs) ? prm.ot[(size_t)(jd0 + lane) * prm.slab_cap + gslab] : 0u;
              for (int jd = 0; jd < lim; ++jd) {
                const uint32_t d = __shfl_sync(0xFFFFFFFFu, d_lane, jd), ow = __shfl_sync(0xFFFFFFFFu, o_lane, jd);
                const uint32_t bit = (ow >> lane) & 1u;
                if (d & 32u) v_hi |= bit << (d & 31u);
                else v_lo |= bit << (d & 31u);
              }
            }
            if (valid) prm.out_rows[r] = (uint64_t)v_lo | ((uint64_t)v_hi << 32);
            continue;
          }
          for (int jd0 = 0; jd0 < n_direct; jd0 += 32) {
            const int lim = min(32, n_direct - jd0);
            const uint32_t fi_lane = lane < lim ? direct_tab[2 * (jd0 + lane)] : 0u;
            const uint32_t dd_lane = lane < lim ? direct_tab[2 * (jd0 + lane) + 1] : 0u;
            for (int jd = 0; jd < lim; ++jd) {
              const uint32_t fi = __shfl_sync(0xFFFFFFFFu, fi_lane, jd), dd = __shfl_sync(0xFFFFFFFFu, dd_lane, jd);
              const uint32_t d = dd & 0x7FFFFFFFu;
              const uint64_t bit = (valid ? ((frow[fi >> 6] >> (fi & 63u)) & 1ull) : 0ull) ^ (uint64_t)(dd >> 31);
              if (d < 64u) v0 |= bit << d;
              else v1 |= bit << (d - 64u);
            }
          }
          for (int jd0 = 0; jd0 < n_draws; jd0 += 32) {
            const int lim = min(32, n_draws - jd0);
            const uint32_t d_lane = lane < lim ? dest[jd0 + lane] : 0u;
            const uint32_t o_lane = (lane < lim && gslab < n_slabs) ? prm.ot[(size_t)(jd0 + lane) * prm.slab_cap + gslab] : 0u;
            for (int jd = 0; jd < lim; ++jd) {
              const uint32_t d = __shfl_sync(0xFFFFFFFFu, d_lane, jd), ow = __shfl_sync(0xFFFFFFFFu, o_lane, jd);
              const uint64_t bit = (uint64_t)((ow >> lane) & 1u);
              if (d < 64u) v0 |= bit << d;
              else v1 |= bit << (d - 64u);
            }
          }
          if (valid) {
            prm.out_rows[r * wo] = v0;
            if (wo > 1) prm.out_rows[r * wo + 1] = v1;
          }
        }
      }
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

// ---------------------------------------------------------------------------------------------
// K1c, parallel form for MODE_FAST companions: all (evaluation, graph) pairs of a component are evaluated at once, one
// thread each, into private accumulators; then one thread per evaluation adds the contributions in graph order -- the
// same additions, in the same order, as the sequential evaluator (adding to a zero accumulator is exact).  The
// component's records are staged in shared memory first when they fit (STAGED), else read through L2.
// One CTA per component.  Dynamic shared memory (words): vals [ne4] | x [n_evals][W] | graph offsets [n_c + 1][max_g] |
// contributions [n_evals][max_g][4] | staged records.
// ---------------------------------------------------------------------------------------------
template <int W, class Src>
__device__ __forceinline__ uint32_t fast_graph_words(const Src& src, uint32_t off) {
  const uint32_t h = src.ld(off);
  const uint32_t nL = h & 0xFFFu, nPi = (h >> 12) & 0xFFFu, nD = h >> 24, nM = src.ld(off + 7);
  return kFastHeaderWords + round4(nL * fast_lin_stride(W)) + round4(nPi * fast_pi_stride(W)) + nM * fast_mpair_stride(W) +
         nD * fast_pair_stride(W);
}

template <int W, class Src>
__device__ __forceinline__ void norm_check_items(const Src& src, const uint32_t* __restrict__ blob, const uint32_t* __restrict__ comp,
                                                 int n_c, int max_g, const uint32_t* __restrict__ xs, uint32_t* __restrict__ goff,
                                                 uint4* __restrict__ contrib, const Tables* tb) {
  const uint32_t* __restrict__ chunk_tab = blob + blob[H_OFF_CHUNK];
  // graph offsets: one thread per level walks the variable-length records
  for (int k = threadIdx.x; k <= n_c; k += blockDim.x) {
    const uint32_t* __restrict__ lvl = blob + blob[H_OFF_LEVEL] + (comp[C_FIRST_LEVEL] + k) * kLevelWords;
    const int first_chunk = (int)lvl[L_FIRST_CHUNK], nck = (int)lvl[L_N_CHUNKS];
    int g = 0;
    for (int c = 0; c < nck; ++c) {
      const uint32_t* row = chunk_tab + (first_chunk + c) * kChunkWords;
      uint32_t off = row[K_OFF];
      for (int i = 0; i < (int)row[K_GRAPHS]; ++i, ++g) {
        goff[k * max_g + g] = off;
        off += fast_graph_words<W>(src, off);
      }
    }
  }
  __syncthreads();
  const int n_evals = 2 * n_c + 1;
  for (int item = threadIdx.x; item < n_evals * max_g; item += blockDim.x) {
    const int t = item / max_g, g = item - t * max_g;
    const int k = (t + 1) >> 1;
    const uint32_t* __restrict__ lvl = blob + blob[H_OFF_LEVEL] + (comp[C_FIRST_LEVEL] + k) * kLevelWords;
    if (g >= (int)lvl[L_G]) continue;
    const bool approx = (lvl[L_FLAGS] & 1u) != 0u;
    uint32_t x[W];
#pragma unroll
    for (int w = 0; w < W; ++w) x[w] = xs[t * W + w];
    LevelAcc acc;
    acc.reset();
    eval_chunk_fast<W>(src, goff[k * max_g + g], 1, approx, x, acc, tb);
    contrib[item] = approx ? make_uint4(__float_as_uint(acc.re), __float_as_uint(acc.im), 0u, 0u)
                           : make_uint4(acc.c.c0, acc.c.c1, acc.c.c2, acc.c.c3);
  }
}

template <int W, bool STAGED>
__global__ void __launch_bounds__(512) norm_check_fast_kernel(const uint32_t* __restrict__ blob, const uint64_t* __restrict__ f_row0,
                                                              const uint64_t* __restrict__ out_row0, float* __restrict__ norm_dev, int max_g) {
  __shared__ Tables tb;
  extern __shared__ __align__(16) float vals[];
  init_tables(&tb, threadIdx.x, blockDim.x);
  const int ci = blockIdx.x;
  const uint32_t* __restrict__ comp = blob + blob[H_OFF_COMP] + ci * kCompWords;
  const int F = (int)comp[C_F], n_c = (int)comp[C_NC], first_draw = (int)comp[C_FIRST_DRAW];
  const uint32_t* __restrict__ sel = blob + blob[H_OFF_FSEL] + comp[C_FSEL_OFF];
  const uint32_t* __restrict__ dest = blob + blob[H_OFF_DEST];
  const int n_evals = 2 * n_c + 1;
  uint32_t* xs = reinterpret_cast<uint32_t*>(vals) + ((n_evals + 3) & ~3);
  uint32_t* goff = xs + ((n_evals * W + 3) & ~3);
  uint4* contrib = reinterpret_cast<uint4*>(goff + (((n_c + 1) * max_g + 3) & ~3));
  // parameter vectors of the evaluations: t = 0 is level 0, t = 2k-1 level k trying bit 1, t = 2k trying bit 0
  for (int i = threadIdx.x; i < n_evals * W; i += blockDim.x) {
    const int t = i / W, w = i - t * W;
    const int k = (t + 1) >> 1;
    const uint32_t trybit = (uint32_t)(t & 1);
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
    if (w == W - 1) xv |= 0x80000000u;  // the always-one parameter of MODE_FAST
    xs[i] = xv;
  }
  const uint32_t* __restrict__ gdata = blob + blob[H_OFF_DATA];
  if constexpr (STAGED) {
    // the component's chunks are contiguous in the data region: [lo, hi)
    const uint32_t* __restrict__ chunk_tab = blob + blob[H_OFF_CHUNK];
    const uint32_t* __restrict__ lv0 = blob + blob[H_OFF_LEVEL] + comp[C_FIRST_LEVEL] * kLevelWords;
    uint32_t lo = 0, hi = 0;
    bool any = false;
    for (int k = 0; k <= n_c; ++k) {
      const uint32_t* lvl = lv0 + k * kLevelWords;
      for (uint32_t c = 0; c < lvl[L_N_CHUNKS]; ++c) {
        const uint32_t* row = chunk_tab + (lvl[L_FIRST_CHUNK] + c) * kChunkWords;
        if (!any) { lo = row[K_OFF]; any = true; }
        hi = row[K_OFF] + row[K_WORDS];
      }
    }
    uint4* staged = contrib + n_evals * max_g;
    const uint4* __restrict__ g4 = reinterpret_cast<const uint4*>(gdata + lo);
    for (uint32_t i = threadIdx.x; i < (hi - lo) / 4; i += blockDim.x) staged[i] = g4[i];
    __syncthreads();
    const SmemSrc src{reinterpret_cast<const uint32_t*>(staged) - lo};
    norm_check_items<W>(src, blob, comp, n_c, max_g, xs, goff, contrib, &tb);
  } else {
    __syncthreads();
    const GmemSrc src{gdata};
    norm_check_items<W>(src, blob, comp, n_c, max_g, xs, goff, contrib, &tb);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_evals; t += blockDim.x) {
    const int k = (t + 1) >> 1;
    const uint32_t* __restrict__ lvl = blob + blob[H_OFF_LEVEL] + (comp[C_FIRST_LEVEL] + k) * kLevelWords;
    const bool approx = (lvl[L_FLAGS] & 1u) != 0u;
    const int G = (int)lvl[L_G];
    LevelAcc acc;
    acc.reset();
    for (int g = 0; g < G; ++g) {
      const uint4 cg = contrib[t * max_g + g];
      if (approx) {
        acc.re = __fadd_rn(acc.re, __uint_as_float(cg.x));
        acc.im = __fadd_rn(acc.im, __uint_as_float(cg.y));
      } else {
        acc.c.c0 += cg.x; acc.c.c1 += cg.y; acc.c.c2 += cg.z; acc.c.c3 += cg.w;
      }
    }
    float re, im;
    finish_level<kModeFast>(acc, approx, (int)lvl[L_P_LO], re, im);
    if (G == 0) { re = 0.0f; im = 0.0f; }
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
