// Post-selection on the device.
//
// The reference buffers survivors on the host (src/tsim/sampler.py:422-545): per chunk of shots it computes the direct
// detector bits from f (`_compute_direct_outputs`, :243-261), discards a shot when a masked direct detector fires
// (:517-520), appends the surviving f rows to a buffer and dispatches sample_program on exactly `batch_size` of them at a
// time (:486-509), scattering the results back by shot index.  Here the same steps run on the GPU, order preserved:
//
//   postselect_flag_kernel     direct bits of every shot -> result row (detector columns), discard flag, survivors per block
//   postselect_scan_kernel     exclusive scan of the block counts on top of the pending count
//   postselect_scatter_kernel  surviving f rows and their shot indices appended to the pending buffer, in shot order
//   pad_rows_kernel            final partial batch: rows n_valid.. = row 0 (fixed batch shape, sampler.py:499-505)
//   scatter_rows_kernel        result[idx[i]] = sampled row i
//   xor_rows_kernel            reference-sample XOR of kept / discarded rows (sampler.py:531-537)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blob.h"

namespace tsb {

constexpr int kPsThreads = 256;

struct PsParams {
  const uint32_t* __restrict__ blob;
  const uint64_t* __restrict__ f;        // [B][wf] chunk rows
  long long B;
  const uint64_t* __restrict__ mask;     // [wo] masked direct detectors (output columns)
  const uint64_t* __restrict__ ref;      // [wo] XORed into the direct bits before the test (zeros: no reference)
  const uint64_t* __restrict__ detmask;  // [wo] columns < num_detectors
  uint64_t* __restrict__ result;         // [B][wo] rows of this chunk inside the result buffer
  uint8_t* __restrict__ discarded;       // [B]
  uint32_t* __restrict__ block_counts;   // [gridDim.x]
};

__device__ __forceinline__ bool ps_direct_row(const uint32_t* __restrict__ blob, const uint64_t* __restrict__ frow, int w,
                                              uint64_t& v) {
  const int n_direct = (int)blob[H_N_DIRECT];
  const uint32_t* __restrict__ direct_tab = blob + blob[H_OFF_DIRECT];
  v = 0;
  for (int j = 0; j < n_direct; ++j) {
    const uint32_t fi = direct_tab[2 * j], dd = direct_tab[2 * j + 1];
    const uint32_t d = dd & 0x7FFFFFFFu;
    if ((int)(d >> 6) != w) continue;
    const uint64_t bit = ((frow[fi >> 6] >> (fi & 63u)) & 1ull) ^ (uint64_t)(dd >> 31);
    v |= bit << (d & 63u);
  }
  return true;
}

__global__ void __launch_bounds__(kPsThreads) postselect_flag_kernel(const PsParams prm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < prm.B;
  const int wf = (int)prm.blob[H_WF64], wo = (int)prm.blob[H_WOUT64];
  bool drop = false;
  if (active) {
    const uint64_t* frow = prm.f + i * wf;
    for (int w = 0; w < wo; ++w) {
      uint64_t v;
      ps_direct_row(prm.blob, frow, w, v);
      if ((v ^ prm.ref[w]) & prm.mask[w]) drop = true;
      prm.result[i * wo + w] = v & prm.detmask[w];
    }
    prm.discarded[i] = drop ? 1 : 0;
  }
  const int n = __syncthreads_count(active && !drop);
  if (threadIdx.x == 0) prm.block_counts[blockIdx.x] = (uint32_t)n;
}

// one CTA: offsets[b] = base + sum_{b' < b} counts[b'];  *total = base + sum of all
__global__ void __launch_bounds__(1024) postselect_scan_kernel(const uint32_t* __restrict__ counts, int n_blocks, uint32_t base,
                                                               uint32_t* __restrict__ offsets, uint32_t* __restrict__ total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = base;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const uint32_t c = b < n_blocks ? counts[b] : 0u;
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      uint32_t s = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    const uint32_t before = carry + (wid ? warp_sums[wid - 1] : 0u) + (x - c);
    if (b < n_blocks) offsets[b] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + c;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kPsThreads) postselect_scatter_kernel(const PsParams prm, const uint32_t* __restrict__ offsets,
                                                                         long long idx_base, uint64_t* __restrict__ surv_f,
                                                                         uint32_t* __restrict__ surv_idx) {
  __shared__ uint32_t warp_base[kPsThreads / 32];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool keep = i < prm.B && prm.discarded[i] == 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
  if (lane == 0) warp_base[wid] = (uint32_t)__popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int k = 0; k < kPsThreads / 32; ++k) {
      const uint32_t c = warp_base[k];
      warp_base[k] = s;
      s += c;
    }
  }
  __syncthreads();
  if (!keep) return;
  const int wf = (int)prm.blob[H_WF64];
  const size_t dst = (size_t)offsets[blockIdx.x] + warp_base[wid] + (uint32_t)__popc(bal & ((1u << lane) - 1u));
  for (int w = 0; w < wf; ++w) surv_f[dst * wf + w] = prm.f[i * wf + w];
  surv_idx[dst] = (uint32_t)(idx_base + i);
}

__global__ void pad_rows_kernel(uint64_t* __restrict__ rows, int words, long long n_valid, long long n_total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (row - n_valid, word)
  if (i >= (n_total - n_valid) * words) return;
  rows[n_valid * words + i] = rows[i % words];
}

__global__ void scatter_rows_kernel(const uint64_t* __restrict__ rows, const uint32_t* __restrict__ idx, long long n, int words,
                                    uint64_t* __restrict__ result) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (row, word)
  if (i >= n * words) return;
  const long long r = i / words;
  const int w = (int)(i % words);
  result[(size_t)idx[r] * words + w] = rows[i];
}

__global__ void xor_rows_kernel(uint64_t* __restrict__ result, const uint8_t* __restrict__ discarded, long long n, int words,
                                const uint64_t* __restrict__ xor_kept, const uint64_t* __restrict__ xor_discarded) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * words) return;
  const long long r = i / words;
  const int w = (int)(i % words);
  result[i] ^= discarded[r] ? xor_discarded[w] : xor_kept[w];
}

}  // namespace tsb
