// libtsim_b200.so -- C ABI (include/tsim_b200.h) + kernels for sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tsim_b200.h"
#include "blob.h"
#include "noise_kernels.cuh"
#include "postselect.cuh"
#include "sampler_kernels.cuh"
#include "sliced_kernels.cuh"
#include <nvtx3/nvToolsExt.h>

namespace tsb {

// =============================================================================================
// K1: sample kernel
// =============================================================================================
//
// dynamic shared memory (32-bit words):
//   [0, 64)                      mbarriers (up to 32 x 8 bytes)
//   [64, 64 + sizeof(Tables)/4)  lookup tables
//   then  wf32 x kThreads        f row of every shot of the tile   (word-major: [w][tid])
//   then  wout32 x kThreads      output row of every shot          (word-major)
//   then  data region (resident) or n_stages x stage_words ring    (128-byte aligned)
// NVTX range around an entry point (header-only NVTX 3: a no-op unless a profiler is attached)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

constexpr int kBarWords = 64;
constexpr int kMaxStages = 32;

template <int W, int MODE>
__global__ void __launch_bounds__(threads_for_words(W), 1) sample_kernel(const KParams prm) {
  constexpr int kThreads = threads_for_words(W);
  extern __shared__ __align__(128) uint32_t smem[];
  const uint32_t* __restrict__ blob = prm.blob;
  const int tid = threadIdx.x;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  Tables* tb = reinterpret_cast<Tables*>(smem + kBarWords);
  const int wf32 = 2 * (int)blob[H_WF64], wout32 = 2 * (int)blob[H_WOUT64];
  uint32_t* sf = smem + kBarWords + sizeof(Tables) / 4;
  uint32_t* so = sf + wf32 * kThreads;
  uint32_t* sdata = smem + prm.smem_data_off;
  const SmemSrc src{sdata};

  const int n_comp = (int)blob[H_N_COMP], n_direct = (int)blob[H_N_DIRECT], n_chunks = (int)blob[H_N_CHUNKS];
  const uint32_t* __restrict__ direct_tab = blob + blob[H_OFF_DIRECT];
  const uint32_t* __restrict__ comp_tab = blob + blob[H_OFF_COMP];
  const uint32_t* __restrict__ level_tab = blob + blob[H_OFF_LEVEL];
  const uint32_t* __restrict__ chunk_tab = blob + blob[H_OFF_CHUNK];
  const uint32_t* __restrict__ fsel = blob + blob[H_OFF_FSEL];
  const uint32_t* __restrict__ dest = blob + blob[H_OFF_DEST];
  const uint32_t* __restrict__ gdata = blob + blob[H_OFF_DATA];

  init_tables(tb, tid, kThreads);
  const int n_bars = prm.resident ? 1 : prm.n_stages;
  if (tid == 0) {
    for (int i = 0; i < n_bars; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();

  // chunk staging.  q = running chunk sequence number of this CTA (tiles x chunks).
  const long long n_rows = prm.rows ? (long long)*prm.n_rows : prm.B;
  const int n_tiles = prm.rows ? (int)((n_rows + kThreads - 1) / kThreads) : prm.n_tiles;
  const int my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total_q = (long long)my_tiles * n_chunks;
  auto issue = [&](long long q) {  // called by thread 0 only
    const int ch = (int)(q % n_chunks);
    const uint32_t* row = chunk_tab + ch * kChunkWords;
    const int stage = (int)(q % prm.n_stages);
    const uint32_t bytes = row[K_WORDS] * 4u;
    mbar_expect_tx(&bars[stage], bytes);
    tma_bulk_g2s(sdata + (size_t)stage * prm.stage_words, gdata + row[K_OFF], bytes, &bars[stage]);
  };
  if (tid == 0 && my_tiles > 0 && n_chunks > 0) {
    if (prm.resident) {
      uint32_t total = blob[H_DATA_WORDS] * 4u;
      mbar_expect_tx(&bars[0], total);
      for (int ch = 0; ch < n_chunks; ++ch) {
        const uint32_t* row = chunk_tab + ch * kChunkWords;
        tma_bulk_g2s(sdata + row[K_OFF], gdata + row[K_OFF], row[K_WORDS] * 4u, &bars[0]);
      }
    } else {
      for (long long q = 0; q < prm.n_stages && q < total_q; ++q) issue(q);
    }
  }
  if (prm.resident && my_tiles > 0 && n_chunks > 0) mbar_wait(&bars[0], 0);

  long long q = 0;  // next chunk sequence number to consume (streaming mode)
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long slot = (long long)tile * kThreads + tid;
    const bool active = slot < n_rows;
    const long long local = (prm.rows && active) ? (long long)prm.rows[slot] : slot;  // row in this launch's buffers
    const unsigned long long shot = (unsigned long long)(prm.shot_offset + local);  // index in the batch (RNG counter)
    const bool is_check = active && shot == 0ull;

    // stage the f row and clear the output row
    for (int w = 0; w < wf32 / 2; ++w) {
      uint64_t v = active ? prm.f[local * (wf32 / 2) + w] : 0ull;
      sf[(2 * w) * kThreads + tid] = (uint32_t)v;
      sf[(2 * w + 1) * kThreads + tid] = (uint32_t)(v >> 32);
    }
    for (int w = 0; w < wout32; ++w) so[w * kThreads + tid] = 0u;

    // direct outputs: f[:, direct_f_indices] ^ direct_flips  (sampler.py:140-145)
    for (int j = 0; j < n_direct; ++j) {
      const uint32_t fi = direct_tab[2 * j], dd = direct_tab[2 * j + 1];
      const uint32_t bit = ((sf[(fi >> 5) * kThreads + tid] >> (fi & 31u)) & 1u) ^ (dd >> 31);
      const uint32_t d = dd & 0x7FFFFFFFu;
      so[(d >> 5) * kThreads + tid] |= bit << (d & 31u);
    }

    for (int ci = 0; ci < n_comp; ++ci) {
      const uint32_t* __restrict__ comp = comp_tab + ci * kCompWords;
      const int F = (int)comp[C_F], n_c = (int)comp[C_NC];
      const uint32_t* __restrict__ sel = fsel + comp[C_FSEL_OFF];
      const int first_draw = (int)comp[C_FIRST_DRAW];
      // gather the selected f bits: f_params[:, f_selection]  (sampler.py:48)
      uint32_t x[W];
#pragma unroll
      for (int w = 0; w < W; ++w) {
        uint32_t xv = 0;
        const int lim = min(32, F - 32 * w);
        for (int b = 0; b < lim; ++b) {
          const uint32_t fi = sel[32 * w + b];
          xv |= ((sf[(fi >> 5) * kThreads + tid] >> (fi & 31u)) & 1u) << b;
        }
        x[w] = xv;
      }
      if (MODE == kModeFast) x[W - 1] |= 0x80000000u;  // always-one parameter

      float prev = 0.0f, dev = 0.0f;
      for (int k = 0; k <= n_c; ++k) {
        const uint32_t* __restrict__ lvl = level_tab + (comp[C_FIRST_LEVEL] + k) * kLevelWords;
        const bool approx = (lvl[L_FLAGS] & 1u) != 0u;
        const int pos = F + k - 1;  // parameter index of the bit being tried
        if (k > 0) {
#pragma unroll
          for (int w = 0; w < W; ++w)
            if (w == (pos >> 5)) x[w] |= 1u << (pos & 31);
        }
        LevelAcc acc, acc0;
        acc.reset();
        acc0.reset();
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
          eval_chunk<W, MODE>(src, off, (int)row[K_GRAPHS], lvl, x, acc, tb);
          if (is_check && k > 0) {
            // normalisation check row: same prefix, trying bit 0 (sampler.py:66-72)
            uint32_t x0[W];
#pragma unroll
            for (int w = 0; w < W; ++w) x0[w] = (w == (pos >> 5)) ? (x[w] & ~(1u << (pos & 31))) : x[w];
            eval_chunk<W, MODE>(src, off, (int)row[K_GRAPHS], lvl, x0, acc0, tb);
          }
          if (!prm.resident) {
            __syncthreads();  // everyone is done with this stage
            if (tid == 0 && q + prm.n_stages < total_q) issue(q + prm.n_stages);
            ++q;
          }
        }
        float re, im;
        finish_level<MODE>(acc, approx, (int)lvl[L_P_LO], re, im);
        if (lvl[L_G] == 0u) { re = 0.0f; im = 0.0f; }  // evaluate.py:34-35
        const float p1 = complex_abs(re, im);
        if (k == 0) {
          prev = p1;
          continue;
        }
        if (is_check) {
          finish_level<MODE>(acc0, approx, (int)lvl[L_P_LO], re, im);
          if (lvl[L_G] == 0u) { re = 0.0f; im = 0.0f; }
          const float p0 = complex_abs(re, im);
          const float norm = __fdiv_rn(__fadd_rn(p0, p1), prev);
          const float d = fabsf(__fsub_rn(norm, 1.0f));
          dev = (dev != dev || d != d) ? __uint_as_float(0x7FC00000u) : fmaxf(dev, d);  // jnp.maximum
        }
        // key, subkey = split(key); bits = bernoulli(subkey, p1 / prev)  (sampler.py:74-75)
        const uint32_t k0 = prm.subkeys[2 * (first_draw + k - 1)], k1 = prm.subkeys[2 * (first_draw + k - 1) + 1];
        const float u = uniform_f32(k0, k1, shot);
        const bool bit = u < __fdiv_rn(p1, prev);
        prev = bit ? p1 : __fsub_rn(prev, p1);  // sampler.py:79
        if (!bit) {
#pragma unroll
          for (int w = 0; w < W; ++w)
            if (w == (pos >> 5)) x[w] &= ~(1u << (pos & 31));
        } else {
          const uint32_t d = dest[first_draw + k - 1];
          so[(d >> 5) * kThreads + tid] |= 1u << (d & 31u);
        }
      }
      if (is_check) prm.norm_dev[ci] = dev;
    }

    if (active) {
      for (int w = 0; w < wout32 / 2; ++w) {
        uint64_t v = (uint64_t)so[(2 * w) * kThreads + tid] | ((uint64_t)so[(2 * w + 1) * kThreads + tid] << 32);
        prm.out[local * (wout32 / 2) + w] = v;
      }
    }
  }
}

// =============================================================================================
// K3: evaluate-only kernel (marginals / probability_of).  Records are read straight from HBM/L2.
// =============================================================================================
template <int W, int MODE>
__global__ void __launch_bounds__(256) evaluate_kernel(const uint32_t* __restrict__ blob, int level_row,
                                                       const uint32_t* __restrict__ xwords, long long B,
                                                       float* __restrict__ amp) {
  __shared__ Tables tb;
  init_tables(&tb, threadIdx.x, blockDim.x);
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const uint32_t* __restrict__ lvl = blob + blob[H_OFF_LEVEL] + level_row * kLevelWords;
  const uint32_t* __restrict__ chunk_tab = blob + blob[H_OFF_CHUNK];
  const GmemSrc src{blob + blob[H_OFF_DATA]};
  uint32_t x[W];
#pragma unroll
  for (int w = 0; w < W; ++w) x[w] = xwords[i * W + w];
  if (MODE == kModeFast) x[W - 1] |= 0x80000000u;
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
  amp[2 * i] = re;
  amp[2 * i + 1] = im;
}

// =============================================================================================
// K0 / K2 helpers: byte rows <-> packed rows
// =============================================================================================
__global__ void pack_rows_kernel(const uint8_t* __restrict__ bytes, long long B, int n_cols, int words64,
                                 uint64_t* __restrict__ packed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (row, word)
  if (i >= B * words64) return;
  const long long row = i / words64;
  const int w = (int)(i % words64);
  const uint8_t* p = bytes + row * n_cols + 64 * w;
  const int lim = min(64, n_cols - 64 * w);
  uint64_t v = 0;
  for (int b = 0; b < lim; ++b) v |= (uint64_t)(p[b] & 1u) << b;
  packed[i] = v;
}

__global__ void unpack_rows_kernel(const uint64_t* __restrict__ packed, long long B, int n_cols, int words64,
                                   uint8_t* __restrict__ bytes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (row, col)
  if (i >= B * n_cols) return;
  const long long row = i / n_cols;
  const int col = (int)(i % n_cols);
  bytes[i] = (uint8_t)((packed[row * words64 + (col >> 6)] >> (col & 63)) & 1ull);
}

// The layout flags of CompiledDetectorSampler.sample (sampler.py:791-868: detector / observable split, prepend / append,
// reference-sample XOR, _maybe_bit_pack) applied to packed rows: up to four column ranges concatenated along the bit axis.
struct LayoutDev {
  int n_seg;
  int lo[4], n[4], start[4];  // source column, length and first output bit of each segment
  int total_bits, row_bytes, bit_packed;
};
__device__ __forceinline__ uint32_t layout_bit(const uint64_t* __restrict__ row, const uint64_t* __restrict__ x, const LayoutDev& L, int t) {
  int col = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < L.n_seg && t >= L.start[i] && t < L.start[i] + L.n[i]) col = L.lo[i] + (t - L.start[i]);
  if (col < 0) return 0u;  // alignment padding between segments
  return (uint32_t)(((row[col >> 6] ^ (x ? x[col >> 6] : 0ull)) >> (col & 63)) & 1ull);
}
// `len` (<= 32) bits of a packed row, starting at column `col`, XOR mask applied
__device__ __forceinline__ uint32_t layout_bits(const uint64_t* __restrict__ row, const uint64_t* __restrict__ x, int col, int len) {
  const int w = col >> 6, off = col & 63;
  uint64_t v = (row[w] ^ (x ? x[w] : 0ull)) >> off;
  if (off + len > 64) v |= (row[w + 1] ^ (x ? x[w + 1] : 0ull)) << (64 - off);
  return (uint32_t)v & (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u));
}
// rows [skip, n_rows) of `packed` -> out rows [0, n_rows - skip).  Bit-packed output: one thread per 32-bit word of an
// output row (word shifts per overlapping column range); bool output: one thread per byte.
__global__ void layout_rows_kernel(const uint64_t* __restrict__ packed, long long n_rows, int skip, int words64,
                                   const uint64_t* __restrict__ xor_row, LayoutDev L, uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows = n_rows - skip;
  if (L.bit_packed) {
    const int wpr = (L.row_bytes + 3) >> 2;  // 32-bit words per output row
    if (i >= rows * wpr) return;
    const long long r = i / wpr;
    const int j = (int)(i % wpr);
    const uint64_t* row = packed + (r + skip) * words64;
    const int t0 = 32 * j;
    uint32_t v = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s >= L.n_seg) break;
      const int a = max(t0, L.start[s]), b = min(t0 + 32, L.start[s] + L.n[s]);
      if (b > a) v |= layout_bits(row, xor_row, L.lo[s] + (a - L.start[s]), b - a) << (a - t0);
    }
    uint8_t* dst = out + r * L.row_bytes + 4 * j;
    const int nb = min(4, L.row_bytes - 4 * j);
    if (nb == 4 && (reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
      *reinterpret_cast<uint32_t*>(dst) = v;
    } else {
      for (int k = 0; k < nb; ++k) dst[k] = (uint8_t)(v >> (8 * k));
    }
    return;
  }
  if (i >= rows * L.row_bytes) return;
  const long long r = i / L.row_bytes;
  const int j = (int)(i % L.row_bytes);
  out[i] = (uint8_t)layout_bit(packed + (r + skip) * words64, xor_row, L, j);
}
// xor_final = xor_in ^ (row0 & ref_mask): the reference sample is in-batch shot 0 (sampler.py:404-409)
__global__ void layout_ref_kernel(const uint64_t* __restrict__ row0, const uint64_t* __restrict__ xor_in, const uint64_t* __restrict__ ref_mask,
                                  int words64, uint64_t* __restrict__ xor_final, uint64_t* __restrict__ row0_copy) {
  const int w = threadIdx.x;
  if (w >= words64) return;
  xor_final[w] = (xor_in ? xor_in[w] : 0ull) ^ (row0[w] & ref_mask[w]);
  row0_copy[w] = row0[w];
}

// draw subkeys on the device: K_0 = batch key; (K_{j+1}, sub_j) = split(K_j)   (sampler.py:74,148)
__global__ void derive_subkeys_kernel(uint32_t k0, uint32_t k1, int n, uint32_t* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int j = 0; j < n; ++j) {
    uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 1;
    threefry2x32(k0, k1, a0, a1);
    threefry2x32(k0, k1, b0, b1);
    k0 = a0; k1 = a1;
    out[2 * j] = b0; out[2 * j + 1] = b1;
  }
}

// params bytes [B, P] -> uint32 [B, W]
__global__ void pack_params_kernel(const uint8_t* __restrict__ bytes, long long B, int P, int W, uint32_t* __restrict__ xw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * W) return;
  const long long row = i / W;
  const int w = (int)(i % W);
  const int lim = min(32, P - 32 * w);
  uint32_t v = 0;
  for (int b = 0; b < lim; ++b) v |= (uint32_t)(bytes[row * P + 32 * w + b] & 1u) << b;
  xw[i] = v;
}


// =============================================================================================
// Pattern cache (optional): shots whose selected f bits have weight <= wmax share their probability
// tree with every other shot of the same pattern.  |E_k(pattern, prefix, 1)| is a pure function of
// (program, pattern, prefix), so it is tabulated once per program with the very same evaluator the
// sampling kernel uses (bit-identical values); a light shot then needs only table walks + its draws.
// Table of one component: [pattern][node], node(k, prefix) = k == 0 ? 0 : 2^(k-1) + prefix.
// Pattern ids (combinatorial number system): 0 = no bit, 1 + i = bit i, 1 + F + C(j,2) + i = bits i < j,
// 1 + F + C(F,2) + C(l,3) + C(j,2) + i = bits i < j < l.
// =============================================================================================
constexpr int kCacheMetaWords = 4;  // per component: table offset (floats), F, n_c, wmax (0xFFFFFFFF = no table)
constexpr int kCacheMaxWords64 = 8;
constexpr int kCacheMaxWeight = 3;

__host__ __device__ inline long long choose2(long long n) { return n * (n - 1) / 2; }
__host__ __device__ inline long long choose3(long long n) { return n * (n - 1) * (n - 2) / 6; }
__host__ __device__ inline long long cache_patterns(long long F, int w) {
  return 1 + (w >= 1 ? F : 0) + (w >= 2 ? choose2(F) : 0) + (w >= 3 ? choose3(F) : 0);
}

template <int W, int MODE>
__global__ void __launch_bounds__(256) cache_build_kernel(const uint32_t* __restrict__ blob, int comp_index, int wmax,
                                                          float* __restrict__ table, long long n_entries) {
  __shared__ Tables tb;
  init_tables(&tb, threadIdx.x, blockDim.x);
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_entries) return;
  const uint32_t* __restrict__ comp = blob + blob[H_OFF_COMP] + comp_index * kCompWords;
  const int F = (int)comp[C_F], n_c = (int)comp[C_NC];
  const int nodes = 1 << n_c;
  const long long pid = t / nodes;
  const int node = (int)(t % nodes);
  const int k = node == 0 ? 0 : 32 - __clz(node);
  const int prefix = node == 0 ? 0 : node - (1 << (k - 1));
  int b0 = -1, b1 = -1, b2 = -1;
  if (pid >= 1 && pid <= F) {
    b0 = pid - 1;
  } else if (pid > F) {
    long long t2 = pid - 1 - F;
    if (t2 >= choose2(F)) {
      t2 -= choose2(F);
      int l = 2;
      while (choose3(l + 1) <= t2) ++l;
      b2 = l;
      t2 -= choose3(l);
    }
    int j = 1;
    while (choose2(j + 1) <= t2) ++j;
    b1 = j;
    b0 = (int)(t2 - choose2(j));
  }
  (void)wmax;
  uint32_t x[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    uint32_t v = 0;
    if (b0 >= 0 && (b0 >> 5) == w) v |= 1u << (b0 & 31);
    if (b1 >= 0 && (b1 >> 5) == w) v |= 1u << (b1 & 31);
    if (b2 >= 0 && (b2 >> 5) == w) v |= 1u << (b2 & 31);
    for (int j = 0; j < k; ++j) {  // prefix bits m_0 .. m_{k-2}, then the trying bit 1
      const int pos = F + j;
      const uint32_t bit = (j == k - 1) ? 1u : (uint32_t)((prefix >> j) & 1);
      if ((pos >> 5) == w) v |= bit << (pos & 31);
    }
    x[w] = v;
  }
  if (MODE == kModeFast) x[W - 1] |= 0x80000000u;
  const uint32_t* __restrict__ lvl = blob + blob[H_OFF_LEVEL] + (comp[C_FIRST_LEVEL] + k) * kLevelWords;
  const uint32_t* __restrict__ chunk_tab = blob + blob[H_OFF_CHUNK];
  const GmemSrc src{blob + blob[H_OFF_DATA]};
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
  table[t] = complex_abs(re, im);
}

struct LightParams {
  const uint32_t* __restrict__ blob;
  const uint64_t* __restrict__ f;
  uint64_t* __restrict__ out;
  const uint32_t* __restrict__ subkeys;
  const float* __restrict__ cache;
  const uint32_t* __restrict__ cache_meta;
  uint32_t* __restrict__ heavy_rows;
  uint32_t* __restrict__ heavy_count;
  long long B;
  long long shot_offset;
  int shot0_heavy;  // per-row programs: in-batch shot 0 carries the norm check, so it takes the full evaluation
};

// pass 1: table walks for light shots; everything else (and shot 0, which carries the norm check) is queued
__global__ void __launch_bounds__(256) light_kernel(const LightParams prm) {
  const uint32_t* __restrict__ blob = prm.blob;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < prm.B;
  const unsigned long long shot = (unsigned long long)(prm.shot_offset + i);
  const int wf = (int)blob[H_WF64], wo = (int)blob[H_WOUT64];
  const int n_comp = (int)blob[H_N_COMP], n_direct = (int)blob[H_N_DIRECT];
  bool heavy = active && shot == 0ull && prm.shot0_heavy;
  uint64_t fw[kCacheMaxWords64], ow[kCacheMaxWords64];
  if (active && !heavy) {
    for (int w = 0; w < wf; ++w) fw[w] = prm.f[i * wf + w];
    for (int w = 0; w < wo; ++w) ow[w] = 0ull;
    const uint32_t* __restrict__ direct_tab = blob + blob[H_OFF_DIRECT];
    for (int j = 0; j < n_direct; ++j) {
      const uint32_t fi = direct_tab[2 * j], dd = direct_tab[2 * j + 1];
      const uint64_t bit = ((fw[fi >> 6] >> (fi & 63u)) & 1ull) ^ (uint64_t)(dd >> 31);
      const uint32_t d = dd & 0x7FFFFFFFu;
      ow[d >> 6] |= bit << (d & 63u);
    }
    const uint32_t* __restrict__ comp_tab = blob + blob[H_OFF_COMP];
    const uint32_t* __restrict__ fsel = blob + blob[H_OFF_FSEL];
    const uint32_t* __restrict__ dest = blob + blob[H_OFF_DEST];
    for (int ci = 0; ci < n_comp && !heavy; ++ci) {
      const uint32_t* __restrict__ comp = comp_tab + ci * kCompWords;
      const uint32_t* __restrict__ meta = prm.cache_meta + ci * kCacheMetaWords;
      const int F = (int)comp[C_F], n_c = (int)comp[C_NC], wmax = (int)meta[3];
      const uint32_t* __restrict__ sel = fsel + comp[C_FSEL_OFF];
      int wgt = 0, p0 = 0, p1 = 0, p2 = 0;
      for (int b = 0; b < F; ++b) {
        const uint32_t fi = sel[b];
        if ((fw[fi >> 6] >> (fi & 63u)) & 1ull) {
          if (wgt == 0) p0 = b;
          else if (wgt == 1) p1 = b;
          else if (wgt == 2) p2 = b;
          ++wgt;
        }
      }
      if (wgt > wmax) { heavy = true; break; }  // wmax == -1: no table for this component
      const long long pid = wgt == 0 ? 0
                          : wgt == 1 ? 1 + p0
                          : wgt == 2 ? 1 + F + choose2(p1) + p0
                                     : 1 + F + choose2(F) + choose3(p2) + choose2(p1) + p0;
      const float* __restrict__ T = prm.cache + meta[0] + ((size_t)pid << n_c);
      float prev = T[0];
      uint32_t prefix = 0;
      const int first_draw = (int)comp[C_FIRST_DRAW];
      for (int k = 1; k <= n_c; ++k) {
        const float p1v = T[(1u << (k - 1)) + prefix];
        const float u = uniform_f32(prm.subkeys[2 * (first_draw + k - 1)], prm.subkeys[2 * (first_draw + k - 1) + 1], shot);
        const bool bit = u < __fdiv_rn(p1v, prev);
        prev = bit ? p1v : __fsub_rn(prev, p1v);
        if (bit) {
          prefix |= 1u << (k - 1);
          const uint32_t d = dest[first_draw + k - 1];
          ow[d >> 6] |= 1ull << (d & 63u);
        }
      }
    }
  }
  // warp-aggregated append of heavy rows
  const unsigned ballot = __ballot_sync(0xFFFFFFFFu, heavy);
  if (ballot) {
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == __ffs(ballot) - 1) base = atomicAdd(prm.heavy_count, (uint32_t)__popc(ballot));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(ballot) - 1);
    if (heavy) prm.heavy_rows[base + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)i;
  }
  if (active && !heavy)
    for (int w = 0; w < wo; ++w) prm.out[i * wo + w] = ow[w];
}

}  // namespace tsb

// =============================================================================================
// host side
// =============================================================================================
using namespace tsb;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                                 \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(TSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Host-side bit packing of the reference's f matrix (uint8[B, num_f], 0/1) into the device's row words.
// The byte matrix is 8x the packed rows: on the 63-byte rows of cfg2 the PCIe copy of the bytes (63 MB per 10^6 shots,
// 1.3 ms at 49 GB/s; 6 ms from pageable memory) is the critical path of the end-to-end call, while a handful of host
// threads pack them at memory speed; only the packed rows then cross the bus.
// ---------------------------------------------------------------------------------------------
class HostPool {
 public:
  explicit HostPool(int n) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // try_begin: the workers start on fn(i), i in [0, n), and the caller carries on; finish: the caller joins in and
  // returns when all are done.  The pool runs one task at a time: try_begin returns false while another thread's task
  // is in flight (several handles driven from several threads, e.g. MultiDeviceProgram) and the caller does its work itself.
  bool try_begin(int n, std::function<void(int)> fn) {
    if (n <= 0 || !busy_.try_lock()) return false;
    {
      std::lock_guard<std::mutex> lk(mu_);
      task_ = std::move(fn);
      fn_ = &task_; next_ = 0; total_ = n; done_ = 0; ++epoch_;
    }
    cv_.notify_all();
    return true;
  }
  void finish() {  // only after a successful try_begin, from the same thread
    work();
    {
      std::unique_lock<std::mutex> lk(mu_);
      done_cv_.wait(lk, [this] { return done_ == total_; });
      fn_ = nullptr;
    }
    busy_.unlock();
  }

 private:
  void work() {
    for (;;) {
      int i;
      const std::function<void(int)>* fn;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!fn_ || next_ >= total_) return;
        i = next_++;
        fn = fn_;
      }
      (*fn)(i);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (++done_ == total_) done_cv_.notify_all();
      }
    }
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
      }
      work();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_, busy_;
  std::condition_variable cv_, done_cv_;
  std::function<void(int)> task_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, done_ = 0;
  unsigned long long epoch_ = 0;
  bool stop_ = false;
};

// threads for host-side packing: TSIM_B200_HOST_THREADS, else the host's cores divided by the ranks sharing them
// (LOCAL_WORLD_SIZE under torchrun), at most 16.  0 or 1 disables the host path.
static int host_threads() {
  static int v = [] {
    if (const char* e = getenv("TSIM_B200_HOST_THREADS")) return std::max(0, atoi(e));
    int hw = (int)std::thread::hardware_concurrency();
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
    return std::max(1, std::min(16, hw / ranks));
  }();
  return v;
}
static HostPool* host_pool() {
  static HostPool* pool = new HostPool(std::max(0, host_threads() - 1));  // the caller works too; leaked on purpose (process lifetime)
  return pool;
}

// SIMD packers with run-time dispatch (host_pack.cpp, plain C++)
extern "C" void tsb_host_pack_rows(const uint8_t* src, long long n, int num_f, int wf, uint64_t* dst, long long avail, int isa);
extern "C" int tsb_host_pack_isa(void);

// rows [0, n) of a uint8[., num_f] matrix -> uint64[., wf] (bit i of a row = f_i & 1), on the pool
// Asynchronous on the pool's workers when the pool is free (returns true: the caller joins with host_pool()->finish());
// otherwise packed right here by the calling thread (returns false).
static bool pack_rows_host_begin(const uint8_t* src, long long n, int num_f, int wf, uint64_t* dst) {
  HostPool* pool = host_pool();
  static const int isa = [] { const char* e = getenv("TSIM_B200_HOST_PACK_ISA"); return e ? atoi(e) : -1; }();
  const int parts = (int)std::min<long long>((pool->size() + 1) * 4, std::max<long long>(1, n / 4096));
  const long long per = (n + parts - 1) / parts;
  if (pool->try_begin(parts, [=](int i) {
        const long long lo = i * per, hi = std::min(n, lo + per);
        if (hi > lo) tsb_host_pack_rows(src + lo * num_f, hi - lo, num_f, wf, dst + lo * wf, (n - lo) * num_f, isa);
      }))
    return true;
  tsb_host_pack_rows(src, n, num_f, wf, dst, n * num_f, isa);
  return false;
}

constexpr int kSlots = 4;            // pipeline depth of tsb_sample_host
constexpr int kHeavyRows = 4;        // word offset of the row list inside a heavy buffer (the count sits at word 0)
static size_t heavy_bytes(long long cap) { return 4 * ((size_t)cap + kHeavyRows + 64); }
constexpr long long kSliceDefault = 196608;  // shots per pipeline slice (tools/e2e_slice_sweep.py: 1.43 / 1.20 ms per 10^6 shots without / with the pattern cache; 262144: 1.45 / 1.27)
static long long slice_env() {
  static long long v = [] {
    if (const char* e = getenv("TSIM_B200_SLICE")) {
      long long x = atoll(e);
      if (x >= 1024) return x;
    }
    return 0ll;
  }();
  return v;
}

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t k_start = nullptr, k_stop = nullptr;
  cudaEvent_t t_h0 = nullptr, t_h1 = nullptr, t_d1 = nullptr;  // TSIM_B200_TRACE: copy-in start / end, copy-out end
  cudaEvent_t e_h2d = nullptr;  // the staging buffer has been copied to the device
  uint64_t* h_stage = nullptr;     // pinned: f rows of the slice packed on the host (host-pack path)
  uint8_t* d_in_bytes = nullptr;   // raw host rows (bytes format) or packed rows
  uint64_t* d_f = nullptr;
  uint64_t* d_out = nullptr;
  uint8_t* d_out_bytes = nullptr;
  uint8_t* d_layout = nullptr;  // rows in the caller's column layout (tsb_sample_noisy_host_layout)
  size_t layout_bytes = 0;
  uint32_t* d_subkeys = nullptr;
  uint32_t* d_heavy = nullptr;  // pattern-cache pass 2: count at [0], row indices from [kHeavyRows] (16-byte aligned)
  uint32_t* d_xt = nullptr;     // sliced mode: transposed parameters [total_F][cap/32]
  uint32_t* d_ot = nullptr;     // sliced mode: bit-sliced outputs [n_draws][cap/32]
  long long cap = 0;
  bool timed = false;
};

struct tsb_program {
  int device = 0;
  std::vector<uint32_t> host_blob;
  uint32_t* d_blob = nullptr;
  float* d_norm_dev = nullptr;
  float* h_norm_dev = nullptr;  // pinned
  uint32_t* d_work = nullptr;
  tsb_info info{};
  int n_stages = 1, stage_words = 0, smem_data_off = 0;
  Slot slots[kSlots];
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  uint32_t* d_subkeys = nullptr;
  float last_ms = 0.f;
  int last_launches = 0;
  int last_sliced_launches = 0;  // kernels of the last launch_sliced call (light pass, K0t, K1s, K2a, norm check)
  int sm_count = 0;
  // pattern cache
  int cache_wmax = -1;  // -1: off
  float* d_cache = nullptr;
  uint32_t* d_cache_meta = nullptr;
  long long cache_entries = 0;
  uint32_t* d_heavy = nullptr;  // for tsb_sample_device
  long long heavy_cap = 0;
  uint32_t* h_heavy_seen = nullptr;  // pinned: heavy-row count of an earlier memoised sliced launch (launch-shape hint only)
  long long heavy_seen_B = 0;
  // MODE_SLICED
  int is_sliced = 0, s_has_exact = 0, s_rows = 0, s_plane_rows = 0, s_smem_limit = 0, s_stage_words = 0, s_index_scale = 1;
  int total_F = 0, max_nc = 0;
  tsb_program* aux = nullptr;   // companion per-row program (norm check); not owned
  cudaStream_t side = nullptr;  // the norm check of shot 0 runs here, overlapped with the rest of the batch
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  uint64_t* d_row0 = nullptr;   // copies of f row 0 and output row 0 for the check
  uint64_t* d_layout_rows = nullptr;  // [4][words_out64]: xor_in | ref_mask | xor_final | copy of row 0 (tsb_sample_noisy_host_layout)
  cudaEvent_t ev_ref = nullptr;       // xor_final is ready
  uint32_t* d_xt = nullptr;     // scratch for tsb_sample_device
  uint32_t* d_ot = nullptr;
  long long scratch_slabs = 0;
};

// Shots per slice of the host pipelines (TSIM_B200_SLICE overrides; TSIM_B200_TRACE=1 prints the timeline).  Measured on
// cfg2 with 10^6 shots: the host-to-device copy of the reference's byte rows (63 MB at ~49 GB/s = 1.3 ms) is the critical
// path, 262144-shot slices leave the shortest tail behind it (tools/sweep_slice.py, tools/e2e_trace.py).
static long long pipeline_slice(const tsb_program*) { return slice_env() ? slice_env() : kSliceDefault; }
// device-noise pipeline (no input copy to hide): 262144-shot slices measured best for cfg2 (0.85 ms per 10^6 shots; 196608:
// 0.98, 524288: 0.83, one slice: 1.07) and cfg3 (8.7 ms per 10^7; 10^6-shot slices: 9.5)
static long long noisy_slice(const tsb_program*) { return slice_env() ? slice_env() : 262144; }

const char* tsb_last_error(void) { return g_err.c_str(); }

int tsb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void tsb_split_key(uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 1;
  threefry2x32(k0, k1, a0, a1);
  threefry2x32(k0, k1, b0, b1);
  out[0] = a0; out[1] = a1; out[2] = b0; out[3] = b1;
}

static int validate_blob(const uint32_t* b, size_t n) {
  if (n < (size_t)kHeaderWords) return fail(TSB_ERR_INVALID, "blob shorter than its header");
  if (b[H_MAGIC] != kMagic) return fail(TSB_ERR_INVALID, "bad magic");
  if (b[H_VERSION] != kVersion) return fail(TSB_ERR_INVALID, "blob version mismatch");
  if (b[H_TOTAL_WORDS] != n) return fail(TSB_ERR_INVALID, "blob size does not match header");
  if (b[H_MODE] > 2u) return fail(TSB_ERR_INVALID, "unknown mode");
  if (b[H_W] < 1u) return fail(TSB_ERR_INVALID, "W must be >= 1");
  const uint32_t offs[] = {b[H_OFF_DIRECT], b[H_OFF_COMP], b[H_OFF_LEVEL], b[H_OFF_CHUNK], b[H_OFF_FSEL], b[H_OFF_DEST], b[H_OFF_DATA]};
  for (uint32_t o : offs)
    if (o > n) return fail(TSB_ERR_INVALID, "table offset out of range");
  if ((size_t)b[H_OFF_DATA] + b[H_DATA_WORDS] > n) return fail(TSB_ERR_INVALID, "data region out of range");
  if (b[H_OFF_DATA] % 4u) return fail(TSB_ERR_INVALID, "data region must be 16-byte aligned");
  const uint32_t* ch = b + b[H_OFF_CHUNK];
  for (uint32_t i = 0; i < b[H_N_CHUNKS]; ++i) {
    if (ch[i * kChunkWords + K_OFF] % 4u || ch[i * kChunkWords + K_WORDS] % 4u)
      return fail(TSB_ERR_INVALID, "chunk not 16-byte aligned");
    if ((size_t)ch[i * kChunkWords + K_OFF] + ch[i * kChunkWords + K_WORDS] > b[H_DATA_WORDS])
      return fail(TSB_ERR_INVALID, "chunk out of range");
  }
  return TSB_OK;
}

// ---- kernel dispatch over (W, MODE) ----------------------------------------------------------
typedef void (*SampleFn)(const KParams);
typedef void (*EvalFn)(const uint32_t*, int, const uint32_t*, long long, float*);

template <int MODE>
static SampleFn sample_fn_for(int W) {
  switch (W) {
    case 1: return sample_kernel<1, MODE>;
    case 2: return sample_kernel<2, MODE>;
    case 3: return sample_kernel<3, MODE>;
    case 4: return sample_kernel<4, MODE>;
    case 5: return sample_kernel<5, MODE>;
    case 6: return sample_kernel<6, MODE>;
    case 7: return sample_kernel<7, MODE>;
    case 8: return sample_kernel<8, MODE>;
    default: return nullptr;
  }
}
template <int MODE>
static EvalFn eval_fn_for(int W) {
  switch (W) {
    case 1: return evaluate_kernel<1, MODE>;
    case 2: return evaluate_kernel<2, MODE>;
    case 3: return evaluate_kernel<3, MODE>;
    case 4: return evaluate_kernel<4, MODE>;
    case 5: return evaluate_kernel<5, MODE>;
    case 6: return evaluate_kernel<6, MODE>;
    case 7: return evaluate_kernel<7, MODE>;
    case 8: return evaluate_kernel<8, MODE>;
    default: return nullptr;
  }
}
typedef void (*SlicedFn)(const SParams);
// exact-branch accumulators are four words per shot: only the 8-way split keeps them in registers; the wide layout
// (64-bit lanes) exists for the 4-way split of all-approximate programs
static SlicedFn sliced_fn(int split, int has_exact, bool rows = false, bool wide = false, bool help = false) {
  if (help) return rows ? sample_sliced_kernel<8, false, true, false, true> : sample_sliced_kernel<8, false, false, false, true>;
  if (wide) return rows ? sample_sliced_kernel<4, false, true, true, false> : sample_sliced_kernel<4, false, false, true, false>;
  if (rows) {
    if (has_exact) return sample_sliced_kernel<8, true, true, false, false>;
    return split == 4 ? sample_sliced_kernel<4, false, true, false, false> : sample_sliced_kernel<8, false, true, false, false>;
  }
  if (has_exact) return sample_sliced_kernel<8, true, false, false, false>;
  return split == 4 ? sample_sliced_kernel<4, false, false, false, false> : sample_sliced_kernel<8, false, false, false, false>;
}

// `ng` = units (32 slabs = 1024 shots) per CTA and round: one narrow group each, or half a wide group (64-bit lanes).
struct SlicedPlan {
  int split = 0, ng = 0, rounds = 0, grid = 0, n_stages = 0, wide = 0, groups = 0, help = 0;
  int xt_off = 0, pl_off = 0, data_off = 0, smem_bytes = 0;
};
// matrix + two plane buffers of one group, in 32-bit words
static int sliced_group_words(int rows, int split, int plane_rows, bool wide = false, bool help = false) {
  return (rows * 32 + 2 * split * (plane_rows + (help ? 1 : 0)) * 32) * (wide ? 2 : 1);  // help: one more plane row per slot
}
// largest number of units (<= cap) that leaves room for `want_stages` stages; 0 if not even one fits
static int sliced_fit_groups(int rows, int plane_rows, int split, int cap, int stage_words, int want_stages, int smem_limit, bool wide = false) {
  for (int nu = cap; nu >= 1; --nu) {
    const int groups = wide ? (nu + 1) / 2 : nu;
    if (((long long)kBarWords + (long long)groups * sliced_group_words(rows, split, plane_rows, wide) + (long long)want_stages * stage_words) * 4 <= smem_limit) return nu;
  }
  return 0;
}
static bool sliced_plan(const tsb_program* p, int n_slabs, SlicedPlan& pl, bool row_list, long long expect_rows);
static SampleFn sample_fn(int mode, int W) { return mode == kModeFast ? sample_fn_for<kModeFast>(W) : sample_fn_for<kModeFaithful>(W); }
static EvalFn eval_fn(int mode, int W) { return mode == kModeFast ? eval_fn_for<kModeFast>(W) : eval_fn_for<kModeFaithful>(W); }

int tsb_program_create(const uint32_t* blob, size_t n_words, int device, tsb_program** out) {
  if (!blob || !out) return fail(TSB_ERR_INVALID, "null argument");
  int rc = validate_blob(blob, n_words);
  if (rc) return rc;
  const int W = (int)blob[H_W], mode = (int)blob[H_MODE];
  if (mode != kModeSliced && !sample_fn(mode, W))
    return fail(TSB_ERR_UNSUPPORTED, "more than 256 parameters per component level (W > 8) is not built");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(TSB_ERR_INVALID, "no such CUDA device");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));

  tsb_program* p = new tsb_program();
  p->device = device;
  p->sm_count = prop.multiProcessorCount;
  p->host_blob.assign(blob, blob + n_words);
  auto bail = [&](int code) { tsb_program_destroy(p); return code; };
#define CUB(call)                                                                                \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      fail(TSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                    \
      return bail(TSB_ERR_CUDA);                                                                 \
    }                                                                                            \
  } while (0)
  CUB(cudaMalloc(&p->d_blob, n_words * 4));
  CUB(cudaMemcpy(p->d_blob, blob, n_words * 4, cudaMemcpyHostToDevice));
  const int n_comp = (int)blob[H_N_COMP], n_draws = (int)blob[H_N_DRAWS];
  CUB(cudaMalloc(&p->d_norm_dev, sizeof(float) * std::max(1, n_comp)));
  CUB(cudaMemset(p->d_norm_dev, 0, sizeof(float) * std::max(1, n_comp)));
  CUB(cudaHostAlloc(&p->h_norm_dev, sizeof(float) * std::max(1, n_comp), cudaHostAllocDefault));
  CUB(cudaHostAlloc(&p->h_heavy_seen, 4, cudaHostAllocDefault));
  *p->h_heavy_seen = 0;
  CUB(cudaMalloc(&p->d_subkeys, 8 * std::max(1, n_draws)));
  CUB(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  CUB(cudaEventCreate(&p->ev_a));
  CUB(cudaEventCreate(&p->ev_b));

  // shared-memory plan
  const int wf32 = 2 * (int)blob[H_WF64], wout32 = 2 * (int)blob[H_WOUT64];
  const int kThreads = threads_for_words(W);
  int fixed_words = kBarWords + (int)(sizeof(Tables) / 4) + (wf32 + wout32) * kThreads;
  if (mode == kModeSliced) {
    p->is_sliced = 1;
    p->s_rows = (int)blob[H_ZERO_ROW] + 1;
    p->s_index_scale = blob[H_INDEX_SCALE] == 2u ? 2 : 1;
    p->s_plane_rows = kMinPlaneRows;
    const uint32_t* lv = blob + blob[H_OFF_LEVEL];
    for (uint32_t i = 0; i < blob[H_N_LEVELS]; ++i)
      if (!(lv[i * kLevelWords + L_FLAGS] & 1u) && lv[i * kLevelWords + L_G] > 0u) p->s_has_exact = 1;
    const uint32_t* ct = blob + blob[H_OFF_COMP];
    for (int c = 0; c < n_comp; ++c) {
      p->total_F += (int)ct[c * kCompWords + C_F];
      p->max_nc = std::max(p->max_nc, (int)ct[c * kCompWords + C_NC]);
    }
    if (p->s_has_exact) p->s_plane_rows = std::max<int>(kMinPlaneRows, (int)blob[H_PLANE_ROWS]);  // multiplied pairs: exact levels only
    CUB(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    CUB(cudaMalloc(&p->d_row0, 8 * (size_t)(blob[H_WF64] + blob[H_WOUT64])));
    int lim_smem = (int)prop.sharedMemPerBlockOptin;
    if (const char* lim = getenv("TSIM_B200_SMEM_LIMIT")) {
      int v = atoi(lim);
      if (v > 0) lim_smem = std::min(lim_smem, v);
    }
    // static shared memory of the kernels (the pair-factor table of the exact variants) comes out of the same budget
    const bool can_wide = !p->s_has_exact && p->s_index_scale == 2;
    for (int split : {4, 8})
      for (bool rows : {false, true})
        for (bool wide : {false, true}) {
          if (wide && !(can_wide && split == 4)) continue;
          cudaFuncAttributes fa;
          CUB(cudaFuncGetAttributes(&fa, (const void*)sliced_fn(split, p->s_has_exact, rows, wide)));
          lim_smem = std::min<int>(lim_smem, (int)prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes);
        }
    if (!p->s_has_exact)
      for (bool rows : {false, true}) {
        cudaFuncAttributes fa;
        CUB(cudaFuncGetAttributes(&fa, (const void*)sliced_fn(8, 0, rows, false, true)));
        lim_smem = std::min<int>(lim_smem, (int)prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes);
      }
    p->s_smem_limit = lim_smem;
    p->s_stage_words = std::max(32, ((int)blob[H_MAX_CHUNK] + 31) & ~31);
    const int split_min = p->s_has_exact ? 8 : 4;
    if (!sliced_fit_groups(p->s_rows, p->s_plane_rows, split_min, 1, p->s_stage_words, 1, lim_smem) ||
        !sliced_fit_groups(p->s_rows, p->s_plane_rows, 8, 1, p->s_stage_words, 1, lim_smem)) {
      fail(TSB_ERR_UNSUPPORTED, "a single chunk of the program does not fit in shared memory next to one slab group");
      return bail(TSB_ERR_UNSUPPORTED);
    }
    for (int split : {4, 8})
      for (bool rows : {false, true})
        for (bool wide : {false, true}) {
          if (wide && !(can_wide && split == 4)) continue;
          CUB(cudaFuncSetAttribute((const void*)sliced_fn(split, p->s_has_exact, rows, wide), cudaFuncAttributeMaxDynamicSharedMemorySize, lim_smem));
        }
    if (!p->s_has_exact)
      for (bool rows : {false, true})
        CUB(cudaFuncSetAttribute((const void*)sliced_fn(8, 0, rows, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize, lim_smem));
    fixed_words = kBarWords;
  }
  fixed_words = (fixed_words + 31) & ~31;  // 128-byte align the data region
  int max_smem = (int)prop.sharedMemPerBlockOptin;
  if (const char* lim = getenv("TSIM_B200_SMEM_LIMIT")) {  // test knob: force the streamed path on small programs
    int v = atoi(lim);
    if (v > 0) max_smem = std::min(max_smem, v);
  }
  const long long budget_words = (long long)max_smem / 4 - fixed_words;
  const long long data_words = blob[H_DATA_WORDS];
  const int max_chunk = (int)blob[H_MAX_CHUNK];
  if (mode != kModeSliced && budget_words < std::max(max_chunk, 4)) {
    fail(TSB_ERR_UNSUPPORTED, "a single chunk of the program does not fit in shared memory");
    return bail(TSB_ERR_UNSUPPORTED);
  }
  int resident = (mode != kModeSliced && data_words <= budget_words) ? 1 : 0;
  long long used;
  if (resident) {
    p->n_stages = 1;
    p->stage_words = (int)data_words;
    used = data_words;
  } else {
    p->stage_words = std::max(32, (max_chunk + 31) & ~31);
    p->n_stages = (int)std::min<long long>(kMaxStages, budget_words / p->stage_words);
    used = (long long)p->n_stages * p->stage_words;
  }
  p->smem_data_off = fixed_words;
  const int smem_bytes = (int)((fixed_words + used) * 4);
  // the attribute is per kernel function and device, not per program: always the device maximum, so that a second
  // (smaller) program of the same (mode, W) cannot lower the limit under an earlier, larger one
  if (mode != kModeSliced)
    CUB(cudaFuncSetAttribute((const void*)sample_fn(mode, W), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));

  tsb_info& in = p->info;
  in.mode = mode; in.words = W; in.num_f = (int)blob[H_NUM_F]; in.num_outputs = (int)blob[H_N_OUT];
  in.n_direct = (int)blob[H_N_DIRECT]; in.n_components = n_comp; in.n_draws = n_draws;
  in.words_f64 = (int)blob[H_WF64]; in.words_out64 = (int)blob[H_WOUT64];
  in.resident = resident; in.n_chunks = (int)blob[H_N_CHUNKS]; in.smem_bytes = smem_bytes;
  in.threads = kThreads; in.grid = p->sm_count; in.data_bytes = data_words * 4;
  if (mode == kModeSliced) {  // report the plan of a chip-filling batch
    SlicedPlan pl;
    sliced_plan(p, p->sm_count * 7 * 32, pl, false, -1);
    in.threads = pl.groups * pl.split * 32; in.smem_bytes = pl.smem_bytes; in.resident = 0;
    p->n_stages = pl.n_stages; p->stage_words = p->s_stage_words;
  }
#undef CUB
  *out = p;
  return TSB_OK;
}

static void free_slot(Slot& s) {
  if (s.h_stage) cudaFreeHost(s.h_stage);
  if (s.d_in_bytes) cudaFree(s.d_in_bytes);
  if (s.d_f) cudaFree(s.d_f);
  if (s.d_out) cudaFree(s.d_out);
  if (s.d_out_bytes) cudaFree(s.d_out_bytes);
  if (s.d_layout) cudaFree(s.d_layout);
  if (s.d_subkeys) cudaFree(s.d_subkeys);
  if (s.d_heavy) cudaFree(s.d_heavy);
  if (s.d_xt) cudaFree(s.d_xt);
  if (s.d_ot) cudaFree(s.d_ot);
  if (s.k_start) cudaEventDestroy(s.k_start);
  if (s.t_h0) cudaEventDestroy(s.t_h0);
  if (s.t_h1) cudaEventDestroy(s.t_h1);
  if (s.t_d1) cudaEventDestroy(s.t_d1);
  if (s.e_h2d) cudaEventDestroy(s.e_h2d);
  if (s.k_stop) cudaEventDestroy(s.k_stop);
  if (s.stream) cudaStreamDestroy(s.stream);
  s = Slot();
}

int tsb_program_destroy(tsb_program* p) {
  if (!p) return TSB_OK;
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  for (auto& s : p->slots) free_slot(s);
  if (p->d_blob) cudaFree(p->d_blob);
  if (p->d_norm_dev) cudaFree(p->d_norm_dev);
  if (p->h_norm_dev) cudaFreeHost(p->h_norm_dev);
  if (p->h_heavy_seen) cudaFreeHost(p->h_heavy_seen);
  if (p->d_subkeys) cudaFree(p->d_subkeys);
  if (p->d_cache) cudaFree(p->d_cache);
  if (p->d_cache_meta) cudaFree(p->d_cache_meta);
  if (p->d_heavy) cudaFree(p->d_heavy);
  if (p->d_xt) cudaFree(p->d_xt);
  if (p->d_ot) cudaFree(p->d_ot);
  if (p->d_row0) cudaFree(p->d_row0);
  if (p->d_layout_rows) cudaFree(p->d_layout_rows);
  if (p->ev_ref) cudaEventDestroy(p->ev_ref);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  if (p->side) cudaStreamDestroy(p->side);
  if (p->ev_a) cudaEventDestroy(p->ev_a);
  if (p->ev_b) cudaEventDestroy(p->ev_b);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return TSB_OK;
}

int tsb_program_info(const tsb_program* p, tsb_info* info) {
  if (!p || !info) return fail(TSB_ERR_INVALID, "null argument");
  *info = p->info;
  return TSB_OK;
}

static int launch_sample_rows(tsb_program* p, const uint64_t* d_f, long long B, long long shot_offset, const uint32_t* d_subkeys,
                              uint64_t* d_out, float* d_norm_dev, cudaStream_t st, const uint32_t* rows, const uint32_t* n_rows) {
  if (B <= 0) return TSB_OK;
  KParams k;
  k.blob = p->d_blob; k.f = d_f; k.out = d_out; k.norm_dev = d_norm_dev; k.subkeys = d_subkeys;
  k.B = B; k.shot_offset = shot_offset;
  const int kThreads = p->info.threads;
  k.n_tiles = (int)((B + kThreads - 1) / kThreads);
  k.resident = p->info.resident; k.n_stages = p->n_stages; k.stage_words = p->stage_words; k.smem_data_off = p->smem_data_off;
  k.rows = rows; k.n_rows = n_rows;
  const int grid = std::min(k.n_tiles, p->info.grid);
  sample_fn(p->info.mode, p->info.words)<<<grid, kThreads, p->info.smem_bytes, st>>>(k);
  CU(cudaGetLastError());
  return TSB_OK;
}

static int launch_light(tsb_program* p, const uint64_t* d_f, long long B, long long shot_offset, const uint32_t* d_subkeys,
                        uint64_t* d_out, uint32_t* heavy_rows, uint32_t* heavy_count, cudaStream_t st);

// One batch slice: either the full evaluation for every shot, or (pattern cache on) table walks for light shots
// followed by the full evaluation of the remaining rows.
static int launch_sample(tsb_program* p, const uint64_t* d_f, long long B, long long shot_offset, const uint32_t* d_subkeys,
                         uint64_t* d_out, float* d_norm_dev, cudaStream_t st, uint32_t* heavy_rows = nullptr,
                         uint32_t* heavy_count = nullptr) {
  if (B <= 0) return TSB_OK;
  if (p->is_sliced) return fail(TSB_ERR_INVALID, "internal: sliced programs go through launch_sliced");
  if (p->cache_wmax < 0 || !heavy_rows) return launch_sample_rows(p, d_f, B, shot_offset, d_subkeys, d_out, d_norm_dev, st, nullptr, nullptr);
  int rc = launch_light(p, d_f, B, shot_offset, d_subkeys, d_out, heavy_rows, heavy_count, st);
  if (rc) return rc;
  return launch_sample_rows(p, d_f, B, shot_offset, d_subkeys, d_out, d_norm_dev, st, heavy_rows, heavy_count);
}


static int launch_light(tsb_program* p, const uint64_t* d_f, long long B, long long shot_offset, const uint32_t* d_subkeys,
                        uint64_t* d_out, uint32_t* heavy_rows, uint32_t* heavy_count, cudaStream_t st) {
  CU(cudaMemsetAsync(heavy_count, 0, 4, st));
  LightParams k;
  k.blob = p->d_blob; k.f = d_f; k.out = d_out; k.subkeys = d_subkeys; k.cache = p->d_cache; k.cache_meta = p->d_cache_meta;
  k.heavy_rows = heavy_rows; k.heavy_count = heavy_count; k.B = B; k.shot_offset = shot_offset;
  k.shot0_heavy = p->is_sliced ? 0 : 1;
  const unsigned blocks = (unsigned)((B + 255) / 256);
  light_kernel<<<blocks, 256, 0, st>>>(k);
  CU(cudaGetLastError());
  return TSB_OK;
}

typedef void (*CacheFn)(const uint32_t*, int, int, float*, long long);
template <int MODE>
static CacheFn cache_fn_for(int W) {
  switch (W) {
    case 1: return cache_build_kernel<1, MODE>;
    case 2: return cache_build_kernel<2, MODE>;
    case 3: return cache_build_kernel<3, MODE>;
    case 4: return cache_build_kernel<4, MODE>;
    case 5: return cache_build_kernel<5, MODE>;
    case 6: return cache_build_kernel<6, MODE>;
    case 7: return cache_build_kernel<7, MODE>;
    case 8: return cache_build_kernel<8, MODE>;
    default: return nullptr;
  }
}

int tsb_program_set_pattern_cache(tsb_program* p, int max_weight, int64_t max_entries, int64_t* entries_out) {
  if (!p) return fail(TSB_ERR_INVALID, "null handle");
  if (max_weight > kCacheMaxWeight) return fail(TSB_ERR_INVALID, "pattern cache supports weights 0 to 3");
  // a sliced program tabulates with its per-row companion's evaluator (bit-identical values, see sliced_kernels.cuh)
  const tsb_program* ev = p->is_sliced ? p->aux : p;
  if (max_weight >= 0 && !ev) return fail(TSB_ERR_UNSUPPORTED, "a sliced program builds its pattern cache with its companion program (tsb_program_set_aux)");
  CU(cudaSetDevice(p->device));
  CU(cudaDeviceSynchronize());
  if (p->d_cache) cudaFree(p->d_cache);
  if (p->d_cache_meta) cudaFree(p->d_cache_meta);
  p->d_cache = nullptr; p->d_cache_meta = nullptr; p->cache_wmax = -1; p->cache_entries = 0;
  if (entries_out) *entries_out = 0;
  if (max_weight < 0) return TSB_OK;
  const tsb_info& in = p->info;
  if (in.words_f64 > kCacheMaxWords64 || in.words_out64 > kCacheMaxWords64)
    return fail(TSB_ERR_UNSUPPORTED, "pattern cache needs num_f <= 512 and num_outputs <= 512");
  if (max_entries <= 0) max_entries = 1ll << 22;  // 16 MB of float32: stays L2-resident
  const uint32_t* b = p->host_blob.data();
  const int n_comp = in.n_components;
  std::vector<uint32_t> meta((size_t)std::max(1, n_comp) * kCacheMetaWords, 0);
  std::vector<long long> offs(n_comp), counts(n_comp);
  long long total = 0;
  for (int c = 0; c < n_comp; ++c) {
    const uint32_t* comp = b + b[H_OFF_COMP] + c * kCompWords;
    const long long F = comp[C_F], n_c = comp[C_NC];
    int w = -1;
    long long cnt = 0;
    if (n_c <= 20) {
      for (int cand = max_weight; cand >= 0; --cand) {
        const long long pats = cache_patterns(F, cand);
        if (pats * (1ll << n_c) <= max_entries - total) { w = cand; cnt = pats * (1ll << n_c); break; }
      }
    }
    offs[c] = total; counts[c] = cnt;
    meta[c * kCacheMetaWords + 0] = (uint32_t)total;
    meta[c * kCacheMetaWords + 1] = (uint32_t)F;
    meta[c * kCacheMetaWords + 2] = (uint32_t)n_c;
    meta[c * kCacheMetaWords + 3] = (uint32_t)w;
    total += cnt;
  }
  if (total >= (1ll << 32)) return fail(TSB_ERR_UNSUPPORTED, "pattern cache too large");
  CU(cudaMalloc(&p->d_cache, sizeof(float) * (size_t)std::max<long long>(1, total)));
  CU(cudaMalloc(&p->d_cache_meta, meta.size() * 4));
  CU(cudaMemcpy(p->d_cache_meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
  CacheFn fn = ev->info.mode == kModeFast ? cache_fn_for<kModeFast>(ev->info.words) : cache_fn_for<kModeFaithful>(ev->info.words);
  for (int c = 0; c < n_comp; ++c) {
    if (counts[c] == 0) continue;
    const unsigned blocks = (unsigned)((counts[c] + 255) / 256);
    fn<<<blocks, 256, 0, p->stream>>>(ev->d_blob, c, (int)meta[c * kCacheMetaWords + 3], p->d_cache + offs[c], counts[c]);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(p->stream));
  p->cache_wmax = max_weight;
  p->cache_entries = total;
  if (entries_out) *entries_out = total;
  return TSB_OK;
}


typedef void (*NormFn)(const uint32_t*, const uint64_t*, const uint64_t*, float*);
template <int MODE>
static NormFn norm_fn_for(int W) {
  switch (W) {
    case 1: return norm_check_kernel<1, MODE>;
    case 2: return norm_check_kernel<2, MODE>;
    case 3: return norm_check_kernel<3, MODE>;
    case 4: return norm_check_kernel<4, MODE>;
    case 5: return norm_check_kernel<5, MODE>;
    case 6: return norm_check_kernel<6, MODE>;
    case 7: return norm_check_kernel<7, MODE>;
    case 8: return norm_check_kernel<8, MODE>;
    default: return nullptr;
  }
}

typedef void (*NormFastFn)(const uint32_t*, const uint64_t*, const uint64_t*, float*, int);
template <bool STAGED>
static NormFastFn norm_fast_fn(int W) {
  switch (W) {
    case 1: return norm_check_fast_kernel<1, STAGED>;
    case 2: return norm_check_fast_kernel<2, STAGED>;
    case 3: return norm_check_fast_kernel<3, STAGED>;
    case 4: return norm_check_fast_kernel<4, STAGED>;
    case 5: return norm_check_fast_kernel<5, STAGED>;
    case 6: return norm_check_fast_kernel<6, STAGED>;
    case 7: return norm_check_fast_kernel<7, STAGED>;
    case 8: return norm_check_fast_kernel<8, STAGED>;
    default: return nullptr;
  }
}

// Launch the parallel norm check on a MODE_FAST companion; false if its scratch does not fit in shared memory.
static bool launch_norm_fast(const tsb_program* p, const tsb_program* a, float* d_norm_dev) {
  const uint32_t* ab = a->host_blob.data();
  const int W = a->info.words;
  int max_g = 1, max_nc = 0;
  long long max_range = 0;
  for (uint32_t i = 0; i < ab[H_N_LEVELS]; ++i) max_g = std::max<int>(max_g, (int)ab[ab[H_OFF_LEVEL] + i * kLevelWords + L_G]);
  for (uint32_t c = 0; c < ab[H_N_COMP]; ++c) {
    const uint32_t* comp = ab + ab[H_OFF_COMP] + c * kCompWords;
    max_nc = std::max<int>(max_nc, (int)comp[C_NC]);
    long long lo = -1, hi = 0;
    for (uint32_t k = 0; k < comp[C_N_LEVELS]; ++k) {
      const uint32_t* lvl = ab + ab[H_OFF_LEVEL] + (comp[C_FIRST_LEVEL] + k) * kLevelWords;
      for (uint32_t j = 0; j < lvl[L_N_CHUNKS]; ++j) {
        const uint32_t* row = ab + ab[H_OFF_CHUNK] + (lvl[L_FIRST_CHUNK] + j) * kChunkWords;
        if (lo < 0) lo = row[K_OFF];
        hi = (long long)row[K_OFF] + row[K_WORDS];
      }
    }
    if (lo >= 0) max_range = std::max(max_range, hi - lo);
  }
  const long long n_evals = 2ll * max_nc + 1;
  const long long base_words = ((n_evals + 3) & ~3ll) + ((n_evals * W + 3) & ~3ll) + (((max_nc + 1ll) * max_g + 3) & ~3ll) + n_evals * max_g * 4;
  const long long limit = p->s_smem_limit - (long long)sizeof(Tables) - 1024;
  const bool staged = (base_words + max_range) * 4 <= limit;
  const long long bytes = (base_words + (staged ? max_range : 0)) * 4;
  if (bytes > limit) return false;
  NormFastFn fn = staged ? norm_fast_fn<true>(W) : norm_fast_fn<false>(W);
  if (!fn) return false;
  // per function and device (not per program): raise it to the most any program may ask for
  if (cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit) != cudaSuccess) return false;
  fn<<<p->info.n_components, 512, (size_t)bytes, p->side>>>(a->d_blob, p->d_row0, p->d_row0 + p->info.words_f64, d_norm_dev, max_g);
  return true;
}

// Groups per CTA, split and ring depth for a batch of n_slabs slabs.  Few groups per SM -> 8 warps per group (so that a
// thin slice still keeps 16+ warps on an SM), otherwise 4.
static bool sliced_plan(const tsb_program* p, int n_slabs, SlicedPlan& pl, bool row_list, long long expect_rows) {
  int n_groups = std::max(1, (n_slabs + 31) / 32);
  pl.grid = std::max(1, std::min(p->sm_count, n_groups));
  int gpc = (n_groups + pl.grid - 1) / pl.grid;  // groups per CTA
  // the smallest grid with that many groups per CTA: the kernel is no slower, and the SMs left over (8 of 148 for 10^6
  // shots) take the neighbouring kernels of the pipeline, the norm check and NCCL's all-gather instead of queueing them
  pl.grid = (n_groups + gpc - 1) / gpc;
  if (row_list) {
    // The live row count is only known on the device: the grid covers the worst case (CTAs beyond the live groups
    // return at once, the kernel derives its rounds itself); the shape (split, groups per CTA) follows the count the
    // previous call saw, which only affects speed.
    pl.grid = std::max(1, std::min(p->sm_count, n_groups));
    if (expect_rows >= 0) {
      const int eg = std::max(1, (int)((expect_rows + 1023) / 1024));
      gpc = (eg + pl.grid - 1) / pl.grid;
    }
  }
  int split = (p->s_has_exact || gpc < 4) ? 8 : 4;
  if (const char* e = getenv("TSIM_B200_SLICED_SPLIT")) {  // tuning knob
    const int v = atoi(e);
    if ((v == 4 && !p->s_has_exact) || v == 8) split = v;
  }
  const int n_chunks = (int)p->host_blob[H_N_CHUNKS];
  const int want = std::min(2, std::max(1, n_chunks));
  // Wide layout (64-bit lanes, two units per group): fewer instructions per shot, fewer warps per SM.  Opt-in through
  // TSIM_B200_SLICED_WIDE=1 (tuning knob).
  bool wide = false;
  int nu_wide = 0;
  if (split == 4 && !p->s_has_exact && p->s_index_scale == 2) {
    nu_wide = sliced_fit_groups(p->s_rows, p->s_plane_rows, 4, std::min(2 * kWideMaxGroups, gpc), p->s_stage_words, want, p->s_smem_limit, true);
    // measured on cfg2 (10^6 shots): 25 % fewer instructions but 16 instead of 28 warps per SM, 0.718 vs 0.687 ms -- the
    // narrow layout stays the default
    if (const char* e = getenv("TSIM_B200_SLICED_WIDE")) wide = atoi(e) != 0 && nu_wide >= 1;
  }
  // Thin launches (one unit per CTA; all-approximate programs): 8 helper warps next to the 8 main warps of the group take
  // the aux part of every graph's stream.  TSIM_B200_SLICED_HELP=0/1 overrides (tuning knob).
  bool help = !wide && split == 8 && !p->s_has_exact && gpc == 1;
  if (const char* e = getenv("TSIM_B200_SLICED_HELP")) help = help && atoi(e) != 0;
  if (help && (kBarWords + sliced_group_words(p->s_rows, p->s_plane_rows, 8, false, true) + (long long)want * p->s_stage_words) * 4 > p->s_smem_limit) help = false;
  int cap = wide ? std::min(2 * kWideMaxGroups, gpc) : std::min(sliced_max_groups(split), gpc);
  int ng = wide ? nu_wide : sliced_fit_groups(p->s_rows, p->s_plane_rows, split, cap, p->s_stage_words, want, p->s_smem_limit);
  if (!ng) ng = sliced_fit_groups(p->s_rows, p->s_plane_rows, split, cap, p->s_stage_words, 1, p->s_smem_limit, wide);
  if (!ng) return false;
  pl.split = split;
  pl.wide = wide ? 1 : 0;
  pl.rounds = (gpc + ng - 1) / ng;
  pl.ng = (gpc + pl.rounds - 1) / pl.rounds;
  if (row_list) pl.ng = ng;  // rounds are derived on the device from the live count
  pl.groups = wide ? (pl.ng + 1) / 2 : pl.ng;
  pl.help = (help && pl.ng == 1) ? 1 : 0;
  const int lane_words = wide ? 2 : 1;
  pl.xt_off = kBarWords;
  pl.pl_off = pl.xt_off + pl.groups * p->s_rows * 32 * lane_words;
  pl.data_off = pl.pl_off + pl.groups * 2 * split * (p->s_plane_rows + pl.help) * 32 * lane_words;
  const long long room = (long long)p->s_smem_limit / 4 - pl.data_off;
  pl.n_stages = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(kSlicedMaxStages, std::max(1, n_chunks)), room / p->s_stage_words));
  pl.smem_bytes = (pl.data_off + pl.n_stages * p->s_stage_words) * 4;
  return true;
}

// [light pass ->] K0t -> K1s -> K2a -> K1c on one stream.  With the pattern cache on (and a heavy buffer), shots whose
// selected f pattern is tabulated are finished by light_kernel; K0t / K1s / K2a then run over the list of remaining rows.
static int launch_sliced(tsb_program* p, const uint64_t* d_f, long long B, long long shot_offset, const uint32_t* d_subkeys,
                         uint64_t* d_out, float* d_norm_dev, cudaStream_t st, uint32_t* d_xt, uint32_t* d_ot, long long slab_cap,
                         bool join = false, cudaEvent_t k1s_start = nullptr, cudaEvent_t k1s_stop = nullptr, uint32_t* d_heavy = nullptr) {
  p->last_sliced_launches = 0;
  if (B <= 0) return TSB_OK;
  const tsb_info& in = p->info;
  const int n_slabs = (int)((B + 31) / 32);
  if (n_slabs > slab_cap) return fail(TSB_ERR_INVALID, "internal: sliced scratch too small");
  const unsigned tblocks = (unsigned)(((long long)n_slabs * 32 + 255) / 256);
  const bool memo = p->cache_wmax >= 0 && d_heavy && in.n_draws > 0;
  const uint32_t* rows = memo ? d_heavy + kHeavyRows : nullptr;
  const uint32_t* n_rows = memo ? d_heavy : nullptr;
  // tsb_last_kernel_ms: the sampling kernel alone, or (memoised) light pass + K0t + K1s + K2a
  if (memo && k1s_start) CU(cudaEventRecord(k1s_start, st));
  if (memo) {
    int rc = launch_light(p, d_f, B, shot_offset, d_subkeys, d_out, d_heavy + kHeavyRows, d_heavy, st);
    if (rc) return rc;
    ++p->last_sliced_launches;
    if (p->h_heavy_seen) {  // remember the count for the next call's launch shape (read whenever it lands)
      CU(cudaMemcpyAsync(p->h_heavy_seen, d_heavy, 4, cudaMemcpyDeviceToHost, st));
      p->heavy_seen_B = B;
    }
  }
  SlicedPlan pl;
  bool fused = false;  // narrow layout: the groups transpose their inputs and assemble their output rows themselves
  if (in.n_draws > 0) {
    long long expect = -1;
    if (memo && p->h_heavy_seen && p->heavy_seen_B > 0)  // scale the last count to this batch size (+ 25 % headroom)
      expect = std::min<long long>(B, (long long)((double)*p->h_heavy_seen * 1.25 * (double)B / (double)p->heavy_seen_B) + 1024);
    if (!sliced_plan(p, n_slabs, pl, memo, expect)) return fail(TSB_ERR_UNSUPPORTED, "internal: no sliced launch plan fits in shared memory");
    fused = !pl.wide && in.words_f64 <= 4 && in.words_out64 <= 2;  // f words held in registers, two output words
    if (const char* e = getenv("TSIM_B200_SLICED_FUSE")) fused = fused && atoi(e) != 0;  // tuning knob: 0 = separate K0t / K2a launches
  }
  if (p->total_F > 0 && in.n_draws > 0 && !fused) {
    transpose_in_kernel<<<tblocks, 256, 0, st>>>(p->d_blob, d_f, B, n_slabs, (int)slab_cap, d_xt, rows, n_rows);
    CU(cudaGetLastError());
    ++p->last_sliced_launches;
  }
  if (in.n_draws > 0) {
    SParams k;
    k.blob = p->d_blob; k.xt = d_xt; k.ot = d_ot; k.subkeys = d_subkeys; k.B = B; k.shot_offset = shot_offset;
    k.pv = reinterpret_cast<float*>(d_ot + (((size_t)slab_cap * std::max(1, in.n_draws) + 3) & ~(size_t)3));  // 16-byte aligned
    k.n_slabs = n_slabs; k.slab_cap = (int)slab_cap;
    k.n_groups = (n_slabs + 31) / 32; k.ng = pl.ng; k.rounds = pl.rounds;
    k.n_stages = pl.n_stages; k.stage_words = p->s_stage_words;
    k.smem_xt_off = pl.xt_off; k.smem_pl_off = pl.pl_off; k.smem_data_off = pl.data_off;
    k.rows = p->s_rows; k.plane_rows = p->s_plane_rows;
    {  // row stride in bytes (128 narrow, 256 wide) over the scale of the stored index bytes
      const uint32_t sb = (pl.wide ? 256u : 128u) / (uint32_t)p->s_index_scale;
      k.sel = make_uint4(sb, sb << 8, sb << 16, sb << 24);
    }
    k.sel_e = make_uint4(8u, 8u << 8, 8u << 16, 8u << 24);
    k.row_list = rows; k.n_rows = n_rows;
    k.f_rows = fused ? d_f : nullptr; k.out_rows = fused ? d_out : nullptr;
    {  // tables behind the ring, if they fit
      const uint32_t* hb = p->host_blob.data();
      const long long tab_words = (long long)hb[H_OFF_FSEL] - (long long)hb[H_OFF_COMP];
      const long long end_words = (long long)pl.data_off + (long long)pl.n_stages * p->s_stage_words;
      k.smem_tab_off = -1; k.tab_words = 0;
      if (tab_words > 0 && tab_words <= 4096 && (end_words + tab_words) * 4 <= p->s_smem_limit) {
        k.smem_tab_off = (int)end_words; k.tab_words = (int)tab_words;
        pl.smem_bytes = (int)((end_words + tab_words) * 4);
      }
    }
    k.lockstep = p->s_has_exact ? 1 : 0;
    if (const char* e = getenv("TSIM_B200_SLICED_LOCKSTEP")) k.lockstep = atoi(e) != 0;  // tuning knob
    if (!memo && k1s_start) CU(cudaEventRecord(k1s_start, st));
    sliced_fn(pl.split, p->s_has_exact, memo, pl.wide != 0, pl.help != 0)<<<pl.grid, pl.groups * pl.split * 32 * (pl.help ? 2 : 1), pl.smem_bytes, st>>>(k);
    CU(cudaGetLastError());
    ++p->last_sliced_launches;
    if (!memo && k1s_stop) CU(cudaEventRecord(k1s_stop, st));
  }
  if (!fused) {
    assemble_out_kernel<<<tblocks, 256, 0, st>>>(p->d_blob, d_f, d_ot, B, n_slabs, (int)slab_cap, d_out, rows, n_rows);
    CU(cudaGetLastError());
    ++p->last_sliced_launches;
  }
  if (memo && k1s_stop) CU(cudaEventRecord(k1s_stop, st));
  if (shot_offset == 0 && in.n_components > 0 && p->aux) {
    ++p->last_sliced_launches;  // the norm check of shot 0 (side stream)
    // fork: the check works on copies of row 0, on a side stream, while st carries on with the next slice / step
    const tsb_program* a = p->aux;
    CU(cudaMemcpyAsync(p->d_row0, d_f, 8 * (size_t)in.words_f64, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(p->d_row0 + in.words_f64, d_out, 8 * (size_t)in.words_out64, cudaMemcpyDeviceToDevice, st));
    CU(cudaEventRecord(p->ev_fork, st));
    CU(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
    NormFn fn = a->info.mode == kModeFast ? norm_fn_for<kModeFast>(a->info.words) : norm_fn_for<kModeFaithful>(a->info.words);
    if (a->info.mode == kModeFast && launch_norm_fast(p, a, d_norm_dev)) fn = nullptr;
    if (fn) fn<<<in.n_components, 128, (2 * p->max_nc + 1) * sizeof(float), p->side>>>(a->d_blob, p->d_row0, p->d_row0 + in.words_f64, d_norm_dev);
    CU(cudaGetLastError());
    CU(cudaEventRecord(p->ev_join, p->side));
    if (join) CU(cudaStreamWaitEvent(st, p->ev_join, 0));
  }
  return TSB_OK;
}

int tsb_program_set_aux(tsb_program* p, tsb_program* aux) {
  if (!p) return fail(TSB_ERR_INVALID, "null handle");
  if (aux) {
    if (aux->is_sliced) return fail(TSB_ERR_INVALID, "the companion program must be a per-row (fast or faithful) program");
    if (aux->device != p->device) return fail(TSB_ERR_INVALID, "companion program lives on another device");
    const tsb_info &a = aux->info, &b = p->info;
    if (a.num_f != b.num_f || a.num_outputs != b.num_outputs || a.n_components != b.n_components || a.n_draws != b.n_draws)
      return fail(TSB_ERR_INVALID, "companion program does not describe the same compiled program");
  }
  p->aux = aux;
  return TSB_OK;
}

int tsb_sample_device(tsb_program* p, const uint64_t* d_f, int64_t B, int64_t shot_offset, uint32_t k0, uint32_t k1,
                      uint64_t* d_out, float* d_norm_dev, void* stream) {
  NvtxRange nvtx_range("tsb_sample_device");
  if (!p) return fail(TSB_ERR_INVALID, "null handle");
  if (B < 0 || shot_offset < 0) return fail(TSB_ERR_INVALID, "negative batch size or offset");
  if (B > 0 && (!d_f || !d_out)) return fail(TSB_ERR_INVALID, "null device buffer");
  CU(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;  // NULL = the default stream, as everywhere in CUDA
  CU(cudaEventRecord(p->ev_a, st));
  if (B > 0) {
    derive_subkeys_kernel<<<1, 32, 0, st>>>(k0, k1, p->info.n_draws, p->d_subkeys);
    CU(cudaGetLastError());
  }
  if (p->is_sliced) {
    const long long slabs = (B + 31) / 32;
    if (p->scratch_slabs < slabs) {
      if (p->d_xt) cudaFree(p->d_xt);
      if (p->d_ot) cudaFree(p->d_ot);
      p->d_xt = nullptr; p->d_ot = nullptr; p->scratch_slabs = 0;
      CU(cudaMalloc(&p->d_xt, 4 * (size_t)slabs * std::max(1, p->total_F)));
      CU(cudaMalloc(&p->d_ot, 4 * ((size_t)slabs * (std::max(1, p->info.n_draws) + 32) + 4)));  // + per-shot chain-rule state
      p->scratch_slabs = slabs;
    }
    // the stream only waits for the (overlapped) norm check when the caller wants the deviations in its own buffer
    if (p->cache_wmax >= 0 && p->heavy_cap < B) {
      if (p->d_heavy) cudaFree(p->d_heavy);
      p->d_heavy = nullptr; p->heavy_cap = 0;
      CU(cudaMalloc(&p->d_heavy, heavy_bytes(B)));
      p->heavy_cap = B;
    }
    int rc = launch_sliced(p, d_f, B, shot_offset, p->d_subkeys, d_out, d_norm_dev ? d_norm_dev : p->d_norm_dev, st, p->d_xt,
                           p->d_ot, p->scratch_slabs, d_norm_dev != nullptr, p->ev_a, p->ev_b, p->cache_wmax >= 0 ? p->d_heavy : nullptr);
    if (rc) return rc;
    if (p->info.n_draws == 0) CU(cudaEventRecord(p->ev_b, st));  // no sampling kernel ran: empty interval
    p->last_launches = B > 0 ? 1 + p->last_sliced_launches : 0;  // derive_subkeys + the sliced pipeline
    p->last_ms = -1.f;
    return TSB_OK;
  }
  if (p->cache_wmax >= 0 && p->heavy_cap < B) {
    if (p->d_heavy) cudaFree(p->d_heavy);
    p->d_heavy = nullptr; p->heavy_cap = 0;
    CU(cudaMalloc(&p->d_heavy, heavy_bytes(B)));
    p->heavy_cap = B;
  }
  int rc = launch_sample(p, d_f, B, shot_offset, p->d_subkeys, d_out, d_norm_dev ? d_norm_dev : p->d_norm_dev, st,
                         p->d_heavy ? p->d_heavy + kHeavyRows : nullptr, p->d_heavy);
  if (rc) return rc;
  CU(cudaEventRecord(p->ev_b, st));
  p->last_launches = B > 0 ? 2 : 0;
  p->last_ms = -1.f;  // resolved lazily in tsb_last_kernel_ms
  return TSB_OK;
}

float tsb_last_kernel_ms(tsb_program* p, int* n_launches) {
  if (!p) return -1.f;
  if (n_launches) *n_launches = p->last_launches;
  if (p->last_ms < 0.f) {
    cudaSetDevice(p->device);
    if (cudaEventSynchronize(p->ev_b) != cudaSuccess) return -1.f;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->ev_a, p->ev_b) != cudaSuccess) return -1.f;
    p->last_ms = ms;
  }
  return p->last_ms;
}

static int ensure_slot(tsb_program* p, Slot& s, long long cap) {
  const tsb_info& in = p->info;
  if (!s.stream) {
    CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s.k_start));
    CU(cudaEventCreate(&s.k_stop));
    CU(cudaEventCreate(&s.t_h0));
    CU(cudaEventCreate(&s.t_h1));
    CU(cudaEventCreate(&s.t_d1));
    CU(cudaEventCreateWithFlags(&s.e_h2d, cudaEventDisableTiming));
    CU(cudaMalloc(&s.d_subkeys, 8 * (size_t)std::max(1, in.n_draws)));
  }
  if (s.cap >= cap) return TSB_OK;
  if (s.h_stage) cudaFreeHost(s.h_stage);
  s.h_stage = nullptr;
  if (s.d_in_bytes) cudaFree(s.d_in_bytes);
  if (s.d_f) cudaFree(s.d_f);
  if (s.d_out) cudaFree(s.d_out);
  if (s.d_out_bytes) cudaFree(s.d_out_bytes);
  if (s.d_heavy) cudaFree(s.d_heavy);
  if (s.d_xt) cudaFree(s.d_xt);
  if (s.d_ot) cudaFree(s.d_ot);
  s.d_in_bytes = nullptr; s.d_f = nullptr; s.d_out = nullptr; s.d_out_bytes = nullptr; s.d_heavy = nullptr;
  s.d_xt = nullptr; s.d_ot = nullptr; s.cap = 0;
  CU(cudaHostAlloc(&s.h_stage, (size_t)cap * in.words_f64 * 8, cudaHostAllocDefault));
  CU(cudaMalloc(&s.d_in_bytes, (size_t)cap * std::max(1, in.num_f)));
  CU(cudaMalloc(&s.d_f, (size_t)cap * in.words_f64 * 8));
  CU(cudaMalloc(&s.d_out, (size_t)cap * in.words_out64 * 8));
  CU(cudaMalloc(&s.d_out_bytes, (size_t)cap * std::max(1, in.num_outputs)));
  CU(cudaMalloc(&s.d_heavy, heavy_bytes(cap)));
  if (p->is_sliced) {
    const size_t slabs = (size_t)((cap + 31) / 32);
    CU(cudaMalloc(&s.d_xt, 4 * slabs * (size_t)std::max(1, p->total_F)));
    CU(cudaMalloc(&s.d_ot, 4 * (slabs * (size_t)(std::max(1, in.n_draws) + 32) + 4)));  // + per-shot chain-rule state
  }
  s.cap = cap;
  return TSB_OK;
}

int tsb_pack_f_device(tsb_program* p, const uint8_t* d_bytes, int64_t B, uint64_t* d_packed, void* stream) {
  if (!p) return fail(TSB_ERR_INVALID, "null handle");
  if (B <= 0) return TSB_OK;
  CU(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * p->info.words_f64;
  if (p->info.num_f == 0) {
    CU(cudaMemsetAsync(d_packed, 0, (size_t)n * 8, st));
    return TSB_OK;
  }
  pack_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_bytes, B, p->info.num_f, p->info.words_f64, d_packed);
  CU(cudaGetLastError());
  return TSB_OK;
}

int tsb_unpack_out_device(tsb_program* p, const uint64_t* d_packed, int64_t B, uint8_t* d_bytes, void* stream) {
  if (!p) return fail(TSB_ERR_INVALID, "null handle");
  if (B <= 0 || p->info.num_outputs == 0) return TSB_OK;
  CU(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * p->info.num_outputs;
  unpack_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_packed, B, p->info.num_outputs, p->info.words_out64, d_bytes);
  CU(cudaGetLastError());
  return TSB_OK;
}

int tsb_sample_host(tsb_program* p, const void* f, int f_format, int64_t B, int64_t shot_offset, uint32_t k0, uint32_t k1,
                    void* out, int out_format, float* norm_dev) {
  NvtxRange nvtx_range("tsb_sample_host");
  if (!p) return fail(TSB_ERR_INVALID, "null handle");
  if (B < 0 || shot_offset < 0) return fail(TSB_ERR_INVALID, "negative batch size or offset");
  if (f_format != TSB_F_BYTES && f_format != TSB_F_PACKED) return fail(TSB_ERR_INVALID, "bad f_format");
  if (out_format != TSB_OUT_BYTES && out_format != TSB_OUT_PACKED) return fail(TSB_ERR_INVALID, "bad out_format");
  const tsb_info& in = p->info;
  if (B > 0 && ((!f && in.num_f > 0) || (!out && in.num_outputs > 0))) return fail(TSB_ERR_INVALID, "null host buffer");
  CU(cudaSetDevice(p->device));
  p->last_ms = 0.f;
  p->last_launches = 0;
  if (norm_dev)
    for (int i = 0; i < in.n_components; ++i) norm_dev[i] = 0.f;
  if (B == 0) return TSB_OK;
  CU(cudaMemsetAsync(p->d_norm_dev, 0, sizeof(float) * std::max(1, in.n_components), p->stream));
  static const bool trace = getenv("TSIM_B200_TRACE") != nullptr;  // per-slice timeline on stderr
  const auto t_call = std::chrono::steady_clock::now();
  if (trace) CU(cudaEventRecord(p->ev_a, p->stream));
  CU(cudaStreamSynchronize(p->stream));
  auto dump = [&](const Slot& s, int idx) {
    float h0 = 0, h1 = 0, k0t = 0, k1t = 0, d1 = 0;
    cudaEventElapsedTime(&h0, p->ev_a, s.t_h0); cudaEventElapsedTime(&h1, p->ev_a, s.t_h1);
    cudaEventElapsedTime(&k0t, p->ev_a, s.k_start); cudaEventElapsedTime(&k1t, p->ev_a, s.k_stop);
    cudaEventElapsedTime(&d1, p->ev_a, s.t_d1);
    fprintf(stderr, "[tsb trace] slice %d: h2d %.3f-%.3f  kernels %.3f-%.3f  d2h done %.3f ms\n", idx, h0, h1, k0t, k1t, d1);
  };

  const long long slice = std::min<long long>(pipeline_slice(p), B);
  const size_t in_row = f_format == TSB_F_BYTES ? (size_t)in.num_f : (size_t)in.words_f64 * 8;
  const size_t out_row = out_format == TSB_OUT_BYTES ? (size_t)in.num_outputs : (size_t)in.words_out64 * 8;
  // byte rows: pack on the host when enough threads are free for this rank (TSIM_B200_HOST_PACK=0/1 overrides)
  const char* host_pack_env = getenv("TSIM_B200_HOST_PACK");
  const bool host_pack = host_pack_env ? atoi(host_pack_env) != 0 : host_threads() >= 6;
  int n_slices = (int)((B + slice - 1) / slice);
  const bool pack_here = f_format == TSB_F_BYTES && host_pack && in.num_f > 0;
  std::chrono::steady_clock::time_point tp0;
  bool pack_async = false;  // the pool's workers are on the current slice (else the calling thread packed it itself)
  struct PoolGuard {  // an error return between begin and join must not leave the pool locked
    bool& on;
    ~PoolGuard() { if (on) host_pool()->finish(); }
  } pool_guard{pack_async};
  // drain(j): the slot of slice j is free again (its previous slice has left the device, results copied out).
  auto drain = [&](int j) -> int {
    Slot& s = p->slots[j % kSlots];
    if (j >= kSlots) {
      CU(cudaStreamSynchronize(s.stream));
      if (trace) dump(s, j - kSlots);
      if (s.timed) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s.k_start, s.k_stop));
        p->last_ms += ms;
        s.timed = false;
      }
    }
    return TSB_OK;
  };
  // stage(j), host packing: as soon as the slot's staging buffer has been copied out, the pool's workers pack slice j
  // into it -- while this thread enqueues the GPU work of slice j - 1
  auto stage = [&](int j) -> int {
    Slot& s = p->slots[j % kSlots];
    int rc = ensure_slot(p, s, slice);
    if (rc) return rc;
    if (!pack_here) return TSB_OK;
    if (j >= kSlots) CU(cudaEventSynchronize(s.e_h2d));
    const long long lo = (long long)j * slice, n = std::min<long long>(slice, B - lo);
    tp0 = std::chrono::steady_clock::now();
    pack_async = pack_rows_host_begin((const uint8_t*)f + (size_t)lo * in_row, n, in.num_f, in.words_f64, s.h_stage);
    return TSB_OK;
  };
  auto pack_join = [&](int j) {
    if (!pack_here) return;
    if (pack_async) host_pool()->finish();
    pack_async = false;
    if (trace)
      fprintf(stderr, "[tsb trace] slice %d: host pack %.3f-%.3f ms (host clock since call start)\n", j,
              std::chrono::duration<double, std::milli>(tp0 - t_call).count(),
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count());
  };
  {
    int rc = stage(0);
    if (rc) return rc;
    pack_join(0);
  }
  for (int i = 0; i < n_slices; ++i) {
    Slot& s = p->slots[i % kSlots];
    int rc = drain(i);
    if (rc) return rc;
    const long long lo = (long long)i * slice, n = std::min<long long>(slice, B - lo);
    derive_subkeys_kernel<<<1, 32, 0, s.stream>>>(k0, k1, in.n_draws, s.d_subkeys);
    CU(cudaGetLastError());
    const uint8_t* src = (const uint8_t*)f + (size_t)lo * in_row;
    if (trace) CU(cudaEventRecord(s.t_h0, s.stream));
    if (pack_here) {
      // packed on the host (pool threads), 1/8 of the bytes cross the bus
      CU(cudaMemcpyAsync(s.d_f, s.h_stage, (size_t)n * in.words_f64 * 8, cudaMemcpyHostToDevice, s.stream));
      CU(cudaEventRecord(s.e_h2d, s.stream));
      if (i + 1 < n_slices) {
        rc = stage(i + 1);
        if (rc) return rc;
      }
    } else if (f_format == TSB_F_BYTES) {
      if (in.num_f > 0) CU(cudaMemcpyAsync(s.d_in_bytes, src, (size_t)n * in_row, cudaMemcpyHostToDevice, s.stream));
      rc = tsb_pack_f_device(p, s.d_in_bytes, n, s.d_f, s.stream);
      if (rc) return rc;
    } else {
      CU(cudaMemcpyAsync(s.d_f, src, (size_t)n * in_row, cudaMemcpyHostToDevice, s.stream));
    }
    if (trace) CU(cudaEventRecord(s.t_h1, s.stream));
    CU(cudaEventRecord(s.k_start, s.stream));
    rc = p->is_sliced ? launch_sliced(p, s.d_f, n, shot_offset + lo, s.d_subkeys, s.d_out, p->d_norm_dev, s.stream, s.d_xt, s.d_ot, (s.cap + 31) / 32, false, nullptr, nullptr, s.d_heavy)
                      : launch_sample(p, s.d_f, n, shot_offset + lo, s.d_subkeys, s.d_out, p->d_norm_dev, s.stream, s.d_heavy + kHeavyRows, s.d_heavy);
    if (rc) return rc;
    CU(cudaEventRecord(s.k_stop, s.stream));
    s.timed = true;
    p->last_launches += 1 + (p->is_sliced ? p->last_sliced_launches : 1) + (f_format == TSB_F_BYTES ? 1 : 0) + (out_format == TSB_OUT_BYTES ? 1 : 0);
    uint8_t* dst = (uint8_t*)out + (size_t)lo * out_row;
    if (out_row > 0) {
      if (out_format == TSB_OUT_BYTES) {
        rc = tsb_unpack_out_device(p, s.d_out, n, s.d_out_bytes, s.stream);
        if (rc) return rc;
        CU(cudaMemcpyAsync(dst, s.d_out_bytes, (size_t)n * out_row, cudaMemcpyDeviceToHost, s.stream));
      } else {
        CU(cudaMemcpyAsync(dst, s.d_out, (size_t)n * out_row, cudaMemcpyDeviceToHost, s.stream));
      }
    }
    if (trace) CU(cudaEventRecord(s.t_d1, s.stream));
    if (i + 1 < n_slices) {
      if (pack_here) {
        pack_join(i + 1);
      } else {
        rc = stage(i + 1);
        if (rc) return rc;
      }
    }
  }
  for (int i = 0; i < kSlots; ++i) {
    Slot& s = p->slots[i];
    if (!s.stream) continue;
    CU(cudaStreamSynchronize(s.stream));
    if (trace && s.timed) dump(s, -1 - i);
    if (s.timed) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, s.k_start, s.k_stop));
      p->last_ms += ms;
      s.timed = false;
    }
  }
  if (p->side) CU(cudaStreamSynchronize(p->side));
  if (norm_dev && in.n_components > 0)
    CU(cudaMemcpy(norm_dev, p->d_norm_dev, sizeof(float) * in.n_components, cudaMemcpyDeviceToHost));
  if (trace)
    fprintf(stderr, "[tsb trace] call returns at %.3f ms (host clock)\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count());
  return TSB_OK;
}

int tsb_evaluate_host(tsb_program* p, int component, int level, const uint8_t* params, int64_t B, float* amp) {
  NvtxRange nvtx_range("tsb_evaluate_host");
  if (!p) return fail(TSB_ERR_INVALID, "null handle");
  if (p->is_sliced) {
    if (!p->aux) return fail(TSB_ERR_UNSUPPORTED, "a sliced program evaluates rows through its companion program (tsb_program_set_aux)");
    return tsb_evaluate_host(p->aux, component, level, params, B, amp);
  }
  const uint32_t* b = p->host_blob.data();
  if (component < 0 || component >= (int)b[H_N_COMP]) return fail(TSB_ERR_INVALID, "component out of range");
  const uint32_t* comp = b + b[H_OFF_COMP] + component * kCompWords;
  if (level < 0 || level >= (int)comp[C_N_LEVELS]) return fail(TSB_ERR_INVALID, "level out of range");
  if (B < 0) return fail(TSB_ERR_INVALID, "negative batch size");
  if (B == 0) return TSB_OK;
  if (!amp) return fail(TSB_ERR_INVALID, "null output");
  const int level_row = (int)comp[C_FIRST_LEVEL] + level;
  const int P = (int)b[b[H_OFF_LEVEL] + level_row * kLevelWords + L_P];
  if (P > 0 && !params) return fail(TSB_ERR_INVALID, "null params");
  const int W = p->info.words;
  CU(cudaSetDevice(p->device));
  uint8_t* d_bytes = nullptr;
  uint32_t* d_x = nullptr;
  float* d_amp = nullptr;
  int rc = TSB_OK;
  cudaError_t e;
#define CUE(call)                                                                 \
  if (rc == TSB_OK && (e = (call)) != cudaSuccess)                                \
    rc = fail(TSB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e));
  CUE(cudaMalloc(&d_bytes, (size_t)B * std::max(1, P)));
  CUE(cudaMalloc(&d_x, (size_t)B * W * 4));
  CUE(cudaMalloc(&d_amp, (size_t)B * 8));
  if (P > 0) CUE(cudaMemcpyAsync(d_bytes, params, (size_t)B * P, cudaMemcpyHostToDevice, p->stream));
  if (rc == TSB_OK) {
    const long long n = (long long)B * W;
    pack_params_kernel<<<(unsigned)((n + 255) / 256), 256, 0, p->stream>>>(d_bytes, B, P, W, d_x);
    eval_fn(p->info.mode, W)<<<(unsigned)((B + 255) / 256), 256, 0, p->stream>>>(p->d_blob, level_row, d_x, B, d_amp);
  }
  CUE(cudaGetLastError());
  CUE(cudaMemcpyAsync(amp, d_amp, (size_t)B * 8, cudaMemcpyDeviceToHost, p->stream));
  CUE(cudaStreamSynchronize(p->stream));
#undef CUE
  if (d_bytes) cudaFree(d_bytes);
  if (d_x) cudaFree(d_x);
  if (d_amp) cudaFree(d_amp);
  return rc;
}

// device-to-device copy on the copy engines (no SM): the push half of the peer-memory gather of output rows
int tsb_memcpy_peer_async(void* dst, const void* src, size_t nbytes, void* stream) {
  if (nbytes == 0) return TSB_OK;
  if (!dst || !src) return fail(TSB_ERR_INVALID, "null pointer");
  CU(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return TSB_OK;
}

void* tsb_host_alloc(size_t nbytes) {
  void* ptr = nullptr;
  if (cudaHostAlloc(&ptr, std::max<size_t>(nbytes, 1), cudaHostAllocDefault) != cudaSuccess) {
    g_err = "cudaHostAlloc failed";
    return nullptr;
  }
  return ptr;
}

void tsb_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

// =============================================================================================
// K5: device channel sampler
// =============================================================================================
struct tsb_noise {
  int device = 0;
  int n_channels = 0, words = 0, n_outcomes = 0;
  uint32_t* d_chan = nullptr;
  uint64_t* d_thr = nullptr;
  uint64_t* d_pat = nullptr;
  double* d_inv = nullptr;  // 1 / log1p(-p_fire) per channel (geometric gaps)
  uint64_t* d_f = nullptr;  // scratch for tsb_noise_sample_host
  long long cap = 0;
  cudaStream_t stream = nullptr;
};

int tsb_noise_create(int n_channels, const int32_t* n_outcomes, const uint64_t* thresholds, const uint64_t* patterns,
                     int words_f64, int device, tsb_noise** out) {
  if (!out || n_channels < 0 || words_f64 < 1) return fail(TSB_ERR_INVALID, "bad argument");
  if (n_channels > 0 && (!n_outcomes || !thresholds || !patterns)) return fail(TSB_ERR_INVALID, "null table");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(TSB_ERR_INVALID, "no such CUDA device");
  CU(cudaSetDevice(device));
  std::vector<uint32_t> chan(2 * (size_t)std::max(1, n_channels));
  std::vector<double> inv((size_t)std::max(1, n_channels), 0.0);
  long long total = 0;
  for (int c = 0; c < n_channels; ++c) {
    if (n_outcomes[c] < 1) return fail(TSB_ERR_INVALID, "a channel needs at least one non-identity outcome");
    chan[2 * c] = (uint32_t)total;
    chan[2 * c + 1] = (uint32_t)n_outcomes[c];
    for (int k = 1; k < n_outcomes[c]; ++k)
      if (thresholds[total + k] < thresholds[total + k - 1]) return fail(TSB_ERR_INVALID, "thresholds must be cumulative");
    total += n_outcomes[c];
    // p_fire = last threshold / 2^64; gaps between fires are geometric: K = floor(ln U / ln(1 - p))
    const double p_fire = std::ldexp((double)thresholds[total - 1], -64);
    inv[c] = p_fire >= 1.0 ? -0.0 : (p_fire > 0.0 ? 1.0 / std::log1p(-p_fire) : -1e300);
  }
  tsb_noise* n = new tsb_noise();
  n->device = device; n->n_channels = n_channels; n->words = words_f64; n->n_outcomes = (int)total;
  cudaError_t e = cudaSuccess;
  auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  chk(cudaMalloc(&n->d_chan, chan.size() * 4));
  chk(cudaMalloc(&n->d_thr, 8 * (size_t)std::max<long long>(1, total)));
  chk(cudaMalloc(&n->d_pat, 8 * (size_t)std::max<long long>(1, total) * words_f64));
  chk(cudaMalloc(&n->d_inv, 8 * inv.size()));
  chk(cudaStreamCreateWithFlags(&n->stream, cudaStreamNonBlocking));
  if (e == cudaSuccess) chk(cudaMemcpy(n->d_chan, chan.data(), chan.size() * 4, cudaMemcpyHostToDevice));
  if (e == cudaSuccess) chk(cudaMemcpy(n->d_inv, inv.data(), inv.size() * 8, cudaMemcpyHostToDevice));
  if (e == cudaSuccess && total > 0) {
    chk(cudaMemcpy(n->d_thr, thresholds, 8 * (size_t)total, cudaMemcpyHostToDevice));
    chk(cudaMemcpy(n->d_pat, patterns, 8 * (size_t)total * words_f64, cudaMemcpyHostToDevice));
  }
  if (e != cudaSuccess) {
    fail(TSB_ERR_CUDA, std::string("tsb_noise_create: ") + cudaGetErrorString(e));
    tsb_noise_destroy(n);
    return TSB_ERR_CUDA;
  }
  *out = n;
  return TSB_OK;
}

int tsb_noise_destroy(tsb_noise* n) {
  if (!n) return TSB_OK;
  cudaSetDevice(n->device);
  if (n->stream) { cudaStreamSynchronize(n->stream); cudaStreamDestroy(n->stream); }
  if (n->d_chan) cudaFree(n->d_chan);
  if (n->d_thr) cudaFree(n->d_thr);
  if (n->d_pat) cudaFree(n->d_pat);
  if (n->d_inv) cudaFree(n->d_inv);
  if (n->d_f) cudaFree(n->d_f);
  delete n;
  return TSB_OK;
}

int tsb_noise_sample_device(tsb_noise* n, int64_t B, int64_t shot_offset, uint64_t seed, uint64_t call, int skip_shot0,
                            uint64_t* d_f, void* stream) {
  NvtxRange nvtx_range("tsb_noise_sample_device");
  if (!n) return fail(TSB_ERR_INVALID, "null handle");
  if (B < 0 || shot_offset < 0) return fail(TSB_ERR_INVALID, "negative batch size or offset");
  if (B == 0) return TSB_OK;
  if (!d_f) return fail(TSB_ERR_INVALID, "null device buffer");
  CU(cudaSetDevice(n->device));
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemsetAsync(d_f, 0, (size_t)B * n->words * 8, st));
  if (n->n_channels == 0) return TSB_OK;
  NoiseParams k;
  k.chan = n->d_chan; k.thresholds = n->d_thr; k.patterns = n->d_pat; k.inv_log1m = n->d_inv; k.f = d_f;
  k.B = B; k.shot_offset = shot_offset; k.n_channels = n->n_channels; k.words = n->words;
  k.seed_lo = (uint32_t)seed; k.seed_hi = (uint32_t)(seed >> 32);
  k.call_lo = (uint32_t)call; k.call_hi = (uint32_t)(call >> 32);
  k.skip_shot0 = skip_shot0;
  k.first_block = shot_offset / kNoiseBlockShots;
  k.n_blocks = (shot_offset + B - 1) / kNoiseBlockShots - k.first_block + 1;
  const long long threads = k.n_blocks * n->n_channels;
  noise_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(k);
  CU(cudaGetLastError());
  return TSB_OK;
}

int tsb_noise_sample_host(tsb_noise* n, int64_t B, int64_t shot_offset, uint64_t seed, uint64_t call, int skip_shot0,
                          uint64_t* f_host) {
  if (!n) return fail(TSB_ERR_INVALID, "null handle");
  if (B < 0) return fail(TSB_ERR_INVALID, "negative batch size");
  if (B == 0) return TSB_OK;
  if (!f_host) return fail(TSB_ERR_INVALID, "null host buffer");
  CU(cudaSetDevice(n->device));
  if (n->cap < B) {
    if (n->d_f) cudaFree(n->d_f);
    n->d_f = nullptr; n->cap = 0;
    CU(cudaMalloc(&n->d_f, (size_t)B * n->words * 8));
    n->cap = B;
  }
  int rc = tsb_noise_sample_device(n, B, shot_offset, seed, call, skip_shot0, n->d_f, n->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(f_host, n->d_f, (size_t)B * n->words * 8, cudaMemcpyDeviceToHost, n->stream));
  CU(cudaStreamSynchronize(n->stream));
  return TSB_OK;
}

// noise -> sample -> (unpack) -> D2H, all on the device: CompiledDetectorSampler.sample() without host noise
// noise -> sample [-> column layout] -> D2H, pipelined over slices; lay == nullptr: whole rows in out_format
static int sample_noisy_host_impl(tsb_program* p, tsb_noise* n, int64_t B, int64_t shot_offset, uint32_t k0, uint32_t k1,
                                  uint64_t noise_seed, uint64_t noise_call, int skip_shot0, void* out, int out_format,
                                  float* norm_dev, uint64_t* f_out, const LayoutDev* lay, const uint64_t* xor_row,
                                  const uint64_t* ref_mask, uint64_t* row0_out, const LayoutDev* lay2 = nullptr, void* out2 = nullptr) {
  NvtxRange nvtx_range("tsb_sample_noisy_host");
  if (!p || !n) return fail(TSB_ERR_INVALID, "null handle");
  if (B < 0 || shot_offset < 0) return fail(TSB_ERR_INVALID, "negative batch size or offset");
  if (!lay && out_format != TSB_OUT_BYTES && out_format != TSB_OUT_PACKED) return fail(TSB_ERR_INVALID, "bad out_format");
  const tsb_info& in = p->info;
  if (n->device != p->device) return fail(TSB_ERR_INVALID, "noise sampler and program live on different devices");
  if (n->words != in.words_f64) return fail(TSB_ERR_INVALID, "noise sampler row width does not match the program");
  if (B > 0 && !out && in.num_outputs > 0) return fail(TSB_ERR_INVALID, "null host buffer");
  if (ref_mask && shot_offset != 0) return fail(TSB_ERR_INVALID, "the reference sample is shot 0 of the batch: shot_offset must be 0");
  CU(cudaSetDevice(p->device));
  p->last_ms = 0.f;
  p->last_launches = 0;
  if (norm_dev)
    for (int i = 0; i < in.n_components; ++i) norm_dev[i] = 0.f;
  if (B == 0) return TSB_OK;
  CU(cudaMemsetAsync(p->d_norm_dev, 0, sizeof(float) * std::max(1, in.n_components), p->stream));
  const int wo = in.words_out64;
  const uint64_t* d_xor = nullptr;
  if (lay && (xor_row || ref_mask)) {
    if (!p->d_layout_rows) {
      CU(cudaMalloc(&p->d_layout_rows, 4 * (size_t)wo * 8));
      CU(cudaEventCreateWithFlags(&p->ev_ref, cudaEventDisableTiming));
    }
    if (xor_row) CU(cudaMemcpyAsync(p->d_layout_rows, xor_row, (size_t)wo * 8, cudaMemcpyHostToDevice, p->stream));
    if (ref_mask) CU(cudaMemcpyAsync(p->d_layout_rows + wo, ref_mask, (size_t)wo * 8, cudaMemcpyHostToDevice, p->stream));
    d_xor = ref_mask ? p->d_layout_rows + 2 * wo : p->d_layout_rows;
  }
  CU(cudaStreamSynchronize(p->stream));
  const long long slice = std::min<long long>(noisy_slice(p), B);
  const size_t out_row = lay ? (size_t)lay->row_bytes : out_format == TSB_OUT_BYTES ? (size_t)in.num_outputs : (size_t)wo * 8;
  const int skip = ref_mask ? 1 : 0;  // the reference row itself is not part of the result (sampler.py:408)
  const int n_slices = (int)((B + slice - 1) / slice);
  for (int i = 0; i < n_slices; ++i) {
    Slot& s = p->slots[i % kSlots];
    int rc = ensure_slot(p, s, slice);
    if (rc) return rc;
    if (i >= kSlots) {
      CU(cudaStreamSynchronize(s.stream));
      if (s.timed) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, s.k_start, s.k_stop));
        p->last_ms += ms;
        s.timed = false;
      }
    }
    const long long lo = (long long)i * slice, cnt = std::min<long long>(slice, B - lo);
    rc = tsb_noise_sample_device(n, cnt, shot_offset + lo, noise_seed, noise_call, skip_shot0, s.d_f, s.stream);
    if (rc) return rc;
    derive_subkeys_kernel<<<1, 32, 0, s.stream>>>(k0, k1, in.n_draws, s.d_subkeys);
    CU(cudaGetLastError());
    CU(cudaEventRecord(s.k_start, s.stream));
    rc = p->is_sliced ? launch_sliced(p, s.d_f, cnt, shot_offset + lo, s.d_subkeys, s.d_out, p->d_norm_dev, s.stream, s.d_xt, s.d_ot, (s.cap + 31) / 32, false, nullptr, nullptr, s.d_heavy)
                      : launch_sample(p, s.d_f, cnt, shot_offset + lo, s.d_subkeys, s.d_out, p->d_norm_dev, s.stream, s.d_heavy + kHeavyRows, s.d_heavy);
    if (rc) return rc;
    CU(cudaEventRecord(s.k_stop, s.stream));
    s.timed = true;
    p->last_launches += 2 + (p->is_sliced ? p->last_sliced_launches : 1) + ((lay || out_format == TSB_OUT_BYTES) ? 1 : 0);  // noise, subkeys, sampling, layout
    if (f_out) CU(cudaMemcpyAsync(f_out + (size_t)lo * in.words_f64, s.d_f, (size_t)cnt * in.words_f64 * 8, cudaMemcpyDeviceToHost, s.stream));
    if (out_row == 0 && !(lay2 && lay2->row_bytes)) continue;
    if (lay) {
      if (ref_mask) {
        if (i == 0) {  // shot 0 of the batch is the reference sample
          layout_ref_kernel<<<1, std::max(32, wo), 0, s.stream>>>(s.d_out, xor_row ? p->d_layout_rows : nullptr, p->d_layout_rows + wo, wo,
                                                                   p->d_layout_rows + 2 * wo, p->d_layout_rows + 3 * wo);
          CU(cudaGetLastError());
          CU(cudaEventRecord(p->ev_ref, s.stream));
          if (row0_out) CU(cudaMemcpyAsync(row0_out, p->d_layout_rows + 3 * wo, (size_t)wo * 8, cudaMemcpyDeviceToHost, s.stream));
        } else {
          CU(cudaStreamWaitEvent(s.stream, p->ev_ref, 0));
        }
      }
      const int sk = i == 0 ? skip : 0;
      const long long rows_out = cnt - sk;
      if (rows_out > 0) {
        const size_t row2 = lay2 ? (size_t)lay2->row_bytes : 0;
        const size_t need = (size_t)slice * (out_row + row2);
        if (s.layout_bytes < need) {
          CU(cudaStreamSynchronize(s.stream));
          if (s.d_layout) cudaFree(s.d_layout);
          s.d_layout = nullptr; s.layout_bytes = 0;
          CU(cudaMalloc(&s.d_layout, need));
          s.layout_bytes = need;
        }
        const size_t first_row = (size_t)(lo - (i == 0 ? 0 : skip));
        const LayoutDev* ls[2] = {lay, lay2};
        uint8_t* hosts[2] = {(uint8_t*)out, (uint8_t*)out2};
        uint8_t* d_dst = s.d_layout;
        for (int a = 0; a < 2; ++a) {  // one or two result arrays (separate_observables)
          if (!ls[a] || ls[a]->row_bytes == 0) continue;
          const size_t rb = (size_t)ls[a]->row_bytes;
          const long long nthreads = rows_out * (long long)(ls[a]->bit_packed ? (rb + 3) / 4 : rb);
          layout_rows_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, s.stream>>>(s.d_out, cnt, sk, wo, d_xor, *ls[a], d_dst);
          CU(cudaGetLastError());
          CU(cudaMemcpyAsync(hosts[a] + first_row * rb, d_dst, (size_t)rows_out * rb, cudaMemcpyDeviceToHost, s.stream));
          d_dst += (size_t)slice * rb;
        }
      }
      continue;
    }
    uint8_t* dst = (uint8_t*)out + (size_t)lo * out_row;
    if (out_format == TSB_OUT_BYTES) {
      rc = tsb_unpack_out_device(p, s.d_out, cnt, s.d_out_bytes, s.stream);
      if (rc) return rc;
      CU(cudaMemcpyAsync(dst, s.d_out_bytes, (size_t)cnt * out_row, cudaMemcpyDeviceToHost, s.stream));
    } else {
      CU(cudaMemcpyAsync(dst, s.d_out, (size_t)cnt * out_row, cudaMemcpyDeviceToHost, s.stream));
    }
  }
  for (int i = 0; i < kSlots; ++i) {
    Slot& s = p->slots[i];
    if (!s.stream) continue;
    CU(cudaStreamSynchronize(s.stream));
    if (s.timed) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, s.k_start, s.k_stop));
      p->last_ms += ms;
      s.timed = false;
    }
  }
  if (p->side) CU(cudaStreamSynchronize(p->side));
  if (norm_dev && in.n_components > 0)
    CU(cudaMemcpy(norm_dev, p->d_norm_dev, sizeof(float) * in.n_components, cudaMemcpyDeviceToHost));
  return TSB_OK;
}

int tsb_sample_noisy_host(tsb_program* p, tsb_noise* n, int64_t B, int64_t shot_offset, uint32_t k0, uint32_t k1,
                          uint64_t noise_seed, uint64_t noise_call, int skip_shot0, void* out, int out_format,
                          float* norm_dev, uint64_t* f_out) {
  return sample_noisy_host_impl(p, n, B, shot_offset, k0, k1, noise_seed, noise_call, skip_shot0, out, out_format, norm_dev, f_out,
                                nullptr, nullptr, nullptr, nullptr);
}

// segments [first, last) of a layout as one result array
static int64_t layout_part_bytes(const tsb_layout* layout, int first, int last) {
  long long bits = 0;
  for (int i = first; i < last; ++i) {
    if (layout->n[i] < 0 || layout->lo[i] < 0) return -1;
    bits += layout->n[i];
  }
  return layout->bit_packed ? (bits + 7) / 8 : bits;
}
static bool layout_ok(const tsb_layout* layout) {
  return layout && layout->n_segments >= 0 && layout->n_segments <= 4 && layout->split >= 0 && layout->split <= layout->n_segments;
}
int tsb_device_mem_info(int device, int64_t* free_bytes, int64_t* total_bytes) {
  size_t f = 0, t = 0;
  CU(cudaSetDevice(device));
  CU(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  return TSB_OK;
}

int64_t tsb_layout_row_bytes(const tsb_layout* layout, int which) {
  if (!layout_ok(layout) || which < 0 || which > 1) return -1;
  const int cut = layout->split > 0 ? layout->split : layout->n_segments;
  return which == 0 ? layout_part_bytes(layout, 0, cut) : layout_part_bytes(layout, cut, layout->n_segments);
}

int tsb_sample_noisy_host_layout(tsb_program* p, tsb_noise* n, int64_t B, int64_t shot_offset, uint32_t k0, uint32_t k1,
                                 uint64_t noise_seed, uint64_t noise_call, int skip_shot0, const tsb_layout* layout,
                                 const uint64_t* xor_row, const uint64_t* ref_mask, uint8_t* out, uint8_t* out2, uint64_t* row0_out,
                                 float* norm_dev) {
  if (!p || !layout) return fail(TSB_ERR_INVALID, "null argument");
  if (!layout_ok(layout) || tsb_layout_row_bytes(layout, 0) < 0 || tsb_layout_row_bytes(layout, 1) < 0)
    return fail(TSB_ERR_INVALID, "bad column layout");
  const int cut = layout->split > 0 ? layout->split : layout->n_segments;
  LayoutDev L[2] = {};
  for (int a = 0; a < 2; ++a) {
    const int first = a == 0 ? 0 : cut, last = a == 0 ? cut : layout->n_segments;
    int bits = 0;
    for (int i = first; i < last; ++i) {
      if (layout->lo[i] + layout->n[i] > p->info.num_outputs) return fail(TSB_ERR_INVALID, "column range exceeds num_outputs");
      const int k = L[a].n_seg++;
      L[a].lo[k] = layout->lo[i]; L[a].n[k] = layout->n[i]; L[a].start[k] = bits;
      bits += layout->n[i];
    }
    L[a].total_bits = bits;
    L[a].bit_packed = layout->bit_packed ? 1 : 0;
    L[a].row_bytes = layout->bit_packed ? (bits + 7) / 8 : bits;
  }
  const bool two = cut < layout->n_segments;
  if (two && B > 0 && L[1].row_bytes > 0 && !out2) return fail(TSB_ERR_INVALID, "null host buffer for the second array");
  return sample_noisy_host_impl(p, n, B, shot_offset, k0, k1, noise_seed, noise_call, skip_shot0, out, TSB_OUT_PACKED, norm_dev, nullptr,
                                &L[0], xor_row, ref_mask, row0_out, two ? &L[1] : nullptr, out2);
}

// =============================================================================================
// Post-selection session (row f3): survivors are compacted, batched, sampled and scattered on the device
// =============================================================================================
struct tsb_postselect {
  tsb_program* p = nullptr;
  long long shots = 0, batch = 0, pushed = 0, pending = 0;
  int wf = 0, wo = 0;
  cudaStream_t stream = nullptr;
  uint64_t* d_rows = nullptr;       // [wo] x 5: mask | ref | detmask | xor_kept | xor_discarded
  uint64_t* d_result = nullptr;     // [shots][wo]
  uint8_t* d_discarded = nullptr;   // [shots]
  uint64_t* d_chunk_f = nullptr;    // [batch][wf]
  uint64_t* d_surv_f = nullptr;     // [2 batch][wf]
  uint32_t* d_surv_idx = nullptr;   // [2 batch]
  uint64_t* d_out = nullptr;        // [batch][wo]
  uint32_t* d_counts = nullptr;     // block counts | block offsets | total
  uint8_t* d_bytes = nullptr;       // [shots][n_out] (finish with byte output)
  float* d_norm = nullptr;
  uint32_t* h_total = nullptr;      // pinned
  float* h_norm = nullptr;          // pinned
  int n_blocks_cap = 0;
  long long dispatches = 0;
};

int tsb_postselect_destroy(tsb_postselect* s) {
  if (!s) return TSB_OK;
  cudaSetDevice(s->p->device);
  if (s->stream) { cudaStreamSynchronize(s->stream); cudaStreamDestroy(s->stream); }
  cudaFree(s->d_rows); cudaFree(s->d_result); cudaFree(s->d_discarded); cudaFree(s->d_chunk_f); cudaFree(s->d_surv_f);
  cudaFree(s->d_surv_idx); cudaFree(s->d_out); cudaFree(s->d_counts); cudaFree(s->d_bytes); cudaFree(s->d_norm);
  if (s->h_total) cudaFreeHost(s->h_total);
  if (s->h_norm) cudaFreeHost(s->h_norm);
  delete s;
  return TSB_OK;
}

int tsb_postselect_create(tsb_program* p, int64_t shots, int64_t batch_size, const uint64_t* mask_row, const uint64_t* ref_row,
                          int num_detectors, tsb_postselect** out) {
  if (!p || !out || !mask_row) return fail(TSB_ERR_INVALID, "null argument");
  if (shots < 0 || batch_size < 1) return fail(TSB_ERR_INVALID, "bad shots / batch_size");
  if (shots >= (1ll << 32)) return fail(TSB_ERR_UNSUPPORTED, "post-selection sessions index shots with 32 bits");
  const tsb_info& in = p->info;
  if (num_detectors < 0 || num_detectors > in.num_outputs) return fail(TSB_ERR_INVALID, "num_detectors out of range");
  CU(cudaSetDevice(p->device));
  tsb_postselect* s = new tsb_postselect();
  s->p = p; s->shots = shots; s->batch = batch_size; s->wf = in.words_f64; s->wo = in.words_out64;
  const int wo = s->wo;
  std::vector<uint64_t> rows(5 * (size_t)wo, 0ull);
  for (int w = 0; w < wo; ++w) {
    rows[w] = mask_row[w];
    rows[wo + w] = ref_row ? ref_row[w] : 0ull;
    const int lo = 64 * w;
    rows[2 * wo + w] = num_detectors >= lo + 64 ? ~0ull : (num_detectors > lo ? ((1ull << (num_detectors - lo)) - 1ull) : 0ull);
  }
  s->n_blocks_cap = (int)((batch_size + kPsThreads - 1) / kPsThreads);
  cudaError_t e = cudaSuccess;
  auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  chk(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  chk(cudaMalloc(&s->d_rows, rows.size() * 8));
  chk(cudaMalloc(&s->d_result, std::max<size_t>(8, (size_t)shots * wo * 8)));
  chk(cudaMalloc(&s->d_discarded, std::max<size_t>(1, (size_t)shots)));
  chk(cudaMalloc(&s->d_chunk_f, (size_t)batch_size * s->wf * 8));
  chk(cudaMalloc(&s->d_surv_f, 2 * (size_t)batch_size * s->wf * 8));
  chk(cudaMalloc(&s->d_surv_idx, 2 * (size_t)batch_size * 4));
  chk(cudaMalloc(&s->d_out, (size_t)batch_size * wo * 8));
  chk(cudaMalloc(&s->d_counts, 4 * (2 * (size_t)s->n_blocks_cap + 4)));
  chk(cudaMalloc(&s->d_norm, sizeof(float) * std::max(1, in.n_components)));
  chk(cudaHostAlloc(&s->h_total, 4, cudaHostAllocDefault));
  chk(cudaHostAlloc(&s->h_norm, sizeof(float) * std::max(1, in.n_components), cudaHostAllocDefault));
  if (e == cudaSuccess) chk(cudaMemcpy(s->d_rows, rows.data(), rows.size() * 8, cudaMemcpyHostToDevice));
  if (e != cudaSuccess) {
    fail(TSB_ERR_CUDA, std::string("tsb_postselect_create: ") + cudaGetErrorString(e));
    tsb_postselect_destroy(s);
    return TSB_ERR_CUDA;
  }
  *out = s;
  return TSB_OK;
}

// the chunk sits in s->d_chunk_f: flag, scan, scatter; returns the pending survivor count
static int postselect_ingest(tsb_postselect* s, long long n, int64_t* pending_out) {
  const tsb_info& in = s->p->info;
  const int wo = s->wo;
  PsParams k;
  k.blob = s->p->d_blob; k.f = s->d_chunk_f; k.B = n;
  k.mask = s->d_rows; k.ref = s->d_rows + wo; k.detmask = s->d_rows + 2 * wo;
  k.result = s->d_result + (size_t)s->pushed * wo; k.discarded = s->d_discarded + s->pushed; k.block_counts = s->d_counts;
  const int nb = (int)((n + kPsThreads - 1) / kPsThreads);
  uint32_t* offsets = s->d_counts + s->n_blocks_cap;
  uint32_t* total = s->d_counts + 2 * s->n_blocks_cap;
  postselect_flag_kernel<<<nb, kPsThreads, 0, s->stream>>>(k);
  postselect_scan_kernel<<<1, 1024, 0, s->stream>>>(s->d_counts, nb, (uint32_t)s->pending, offsets, total);
  postselect_scatter_kernel<<<nb, kPsThreads, 0, s->stream>>>(k, offsets, s->pushed, s->d_surv_f, s->d_surv_idx);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(s->h_total, total, 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->pending = (long long)*s->h_total;
  s->pushed += n;
  (void)in;
  if (pending_out) *pending_out = s->pending;
  return TSB_OK;
}

static int postselect_check_push(tsb_postselect* s, int64_t n) {
  if (!s) return fail(TSB_ERR_INVALID, "null session");
  if (n < 0 || n > s->batch) return fail(TSB_ERR_INVALID, "a chunk holds at most batch_size shots");
  if (s->pushed + n > s->shots) return fail(TSB_ERR_INVALID, "more shots pushed than the session was created for");
  if (s->pending >= s->batch) return fail(TSB_ERR_INVALID, "dispatch the pending full batch before pushing more shots");
  return TSB_OK;
}

int tsb_postselect_push_host(tsb_postselect* s, const uint64_t* f_packed, int64_t n, int64_t* pending_out) {
  int rc = postselect_check_push(s, n);
  if (rc) return rc;
  if (n == 0) { if (pending_out) *pending_out = s->pending; return TSB_OK; }
  if (!f_packed) return fail(TSB_ERR_INVALID, "null f rows");
  CU(cudaSetDevice(s->p->device));
  CU(cudaMemcpyAsync(s->d_chunk_f, f_packed, (size_t)n * s->wf * 8, cudaMemcpyHostToDevice, s->stream));
  return postselect_ingest(s, n, pending_out);
}

int tsb_postselect_push_noise(tsb_postselect* s, tsb_noise* noise, int64_t n, uint64_t seed, uint64_t call, int64_t* pending_out) {
  int rc = postselect_check_push(s, n);
  if (rc) return rc;
  if (n == 0) { if (pending_out) *pending_out = s->pending; return TSB_OK; }
  if (!noise || noise->device != s->p->device || noise->words != s->wf) return fail(TSB_ERR_INVALID, "noise sampler does not match the program");
  CU(cudaSetDevice(s->p->device));
  rc = tsb_noise_sample_device(noise, n, 0, seed, call, 0, s->d_chunk_f, s->stream);
  if (rc) return rc;
  return postselect_ingest(s, n, pending_out);
}

int tsb_postselect_dispatch(tsb_postselect* s, uint32_t k0, uint32_t k1, int final_batch, float* norm_dev, int64_t* pending_out) {
  NvtxRange nvtx_range("tsb_postselect_dispatch");
  if (!s) return fail(TSB_ERR_INVALID, "null session");
  const tsb_info& in = s->p->info;
  if (s->pending == 0) return fail(TSB_ERR_INVALID, "no pending survivors");
  if (s->pending < s->batch && !final_batch) return fail(TSB_ERR_INVALID, "a partial batch is only dispatched at the end");
  CU(cudaSetDevice(s->p->device));
  const long long n_valid = std::min(s->pending, s->batch);
  if (n_valid < s->batch) {  // fixed batch shape: pad with copies of the first pending row (sampler.py:499-505)
    const long long n = (s->batch - n_valid) * s->wf;
    pad_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->d_surv_f, s->wf, n_valid, s->batch);
    CU(cudaGetLastError());
  }
  CU(cudaMemsetAsync(s->d_norm, 0, sizeof(float) * std::max(1, in.n_components), s->stream));
  int rc = tsb_sample_device(s->p, s->d_surv_f, s->batch, 0, k0, k1, s->d_out, s->d_norm, s->stream);
  if (rc) return rc;
  {
    const long long n = n_valid * s->wo;
    scatter_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->d_out, s->d_surv_idx, n_valid, s->wo, s->d_result);
    CU(cudaGetLastError());
  }
  const long long left = s->pending - n_valid;
  if (left > 0) {  // left < batch: source [batch, batch + left) and destination [0, left) do not overlap
    CU(cudaMemcpyAsync(s->d_surv_f, s->d_surv_f + (size_t)s->batch * s->wf, (size_t)left * s->wf * 8, cudaMemcpyDeviceToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_surv_idx, s->d_surv_idx + s->batch, (size_t)left * 4, cudaMemcpyDeviceToDevice, s->stream));
  }
  if (in.n_components > 0)
    CU(cudaMemcpyAsync(s->h_norm, s->d_norm, sizeof(float) * in.n_components, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (norm_dev)
    for (int i = 0; i < in.n_components; ++i) norm_dev[i] = s->h_norm[i];
  s->pending = left;
  s->dispatches += 1;
  if (pending_out) *pending_out = left;
  return TSB_OK;
}

int tsb_postselect_finish(tsb_postselect* s, const uint64_t* xor_kept, const uint64_t* xor_discarded, void* out, int out_format,
                          uint8_t* discarded_out) {
  if (!s) return fail(TSB_ERR_INVALID, "null session");
  if (s->pushed != s->shots) return fail(TSB_ERR_INVALID, "not every shot of the session was pushed");
  if (s->pending != 0) return fail(TSB_ERR_INVALID, "pending survivors were not dispatched");
  if (out_format != TSB_OUT_BYTES && out_format != TSB_OUT_PACKED) return fail(TSB_ERR_INVALID, "bad out_format");
  if (s->shots == 0) return TSB_OK;
  const tsb_info& in = s->p->info;
  CU(cudaSetDevice(s->p->device));
  const int wo = s->wo;
  if (xor_kept && xor_discarded) {
    CU(cudaMemcpyAsync(s->d_rows + 3 * wo, xor_kept, 8 * (size_t)wo, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(s->d_rows + 4 * wo, xor_discarded, 8 * (size_t)wo, cudaMemcpyHostToDevice, s->stream));
    const long long n = s->shots * wo;
    xor_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->d_result, s->d_discarded, s->shots, wo, s->d_rows + 3 * wo, s->d_rows + 4 * wo);
    CU(cudaGetLastError());
  }
  if (out && in.num_outputs > 0) {
    if (out_format == TSB_OUT_BYTES) {
      if (!s->d_bytes) CU(cudaMalloc(&s->d_bytes, (size_t)s->shots * in.num_outputs));
      int rc = tsb_unpack_out_device(s->p, s->d_result, s->shots, s->d_bytes, s->stream);
      if (rc) return rc;
      CU(cudaMemcpyAsync(out, s->d_bytes, (size_t)s->shots * in.num_outputs, cudaMemcpyDeviceToHost, s->stream));
    } else {
      CU(cudaMemcpyAsync(out, s->d_result, (size_t)s->shots * wo * 8, cudaMemcpyDeviceToHost, s->stream));
    }
  }
  if (discarded_out) CU(cudaMemcpyAsync(discarded_out, s->d_discarded, (size_t)s->shots, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return TSB_OK;
}
