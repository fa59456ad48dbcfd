// K5: error-mechanism (channel) sampler on the device.
//
// Replaces the host loop of the reference's ChannelSampler.sample (src/tsim/noise/channels.py:624-658):
// every channel fires in a shot with probability p_fire; a fired channel picks a non-identity outcome
// from its conditional distribution and XORs that outcome's precomputed f-pattern into the shot's row.
// The reference draws geometric skips from NumPy's PCG64 stream (channels.py:638-656), which cannot be reproduced in
// parallel, so parity here is statistical (same distribution, different stream).  Like the reference the kernel walks
// from fire to fire: a thread owns one channel over one block of kNoiseBlockShots in-batch shots and draws geometric
// gaps K = floor(ln U / ln(1 - p_fire)) (P(K >= k) = (1 - p)^k) from a counter-based Philox4x32-10 generator keyed by
// (seed, call) with counter (block, channel, draw) -- work is O(fires + channels * blocks), not O(shots * channels), and
// the f rows stay a pure function of (seed, call, in-batch shot index, channel): a shard applies the fires that fall
// inside its row range, so rows are independent of batch partitioning and of the number of GPUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsb {

struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ ka; c1 = lo1; c2 = hi0 ^ c3 ^ kb; c3 = lo0;
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

constexpr int kNoiseBlockShots = 1024;

struct NoiseParams {
  const uint32_t* __restrict__ chan;        // [n_channels][2]: first outcome index, number of non-identity outcomes
  const uint64_t* __restrict__ thresholds;  // [n_outcomes_total] cumulative, scaled to 2^64: p_fire = thresholds[last of channel] / 2^64
  const uint64_t* __restrict__ patterns;    // [n_outcomes_total][words]
  const double* __restrict__ inv_log1m;     // [n_channels]: 1 / log1p(-p_fire) (<= 0; -0 for p_fire = 1: every shot fires)
  uint64_t* __restrict__ f;                 // [B][words], zero-initialised by the caller
  long long B;
  long long shot_offset;
  long long first_block, n_blocks;  // blocks of kNoiseBlockShots in-batch shots that overlap [shot_offset, shot_offset + B)
  int n_channels;
  int words;
  uint32_t seed_lo, seed_hi;
  uint32_t call_lo, call_hi;
  int skip_shot0;       // leave in-batch shot 0 noiseless (reference-sample row, sampler.py:395-396)
};

// one thread per (channel, block of shots); neighbouring threads share the channel (similar trip counts)
__global__ void __launch_bounds__(128) noise_kernel(const NoiseParams prm) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= prm.n_blocks * prm.n_channels) return;
  const int c = (int)(idx / prm.n_blocks);
  const unsigned long long blk = (unsigned long long)(prm.first_block + idx % prm.n_blocks);
  const uint32_t first = prm.chan[2 * c], m = prm.chan[2 * c + 1];
  const uint64_t t_last = prm.thresholds[first + m - 1];
  if (t_last == 0ull) return;  // never fires
  const double inv = prm.inv_log1m[c];
  const Philox rng{prm.seed_lo ^ prm.call_lo * 0x9E3779B9u, prm.seed_hi ^ prm.call_hi};
  const long long lo = prm.shot_offset, hi = prm.shot_offset + prm.B;
  long long pos = -1;
  for (uint32_t draw = 0;; ++draw) {
    uint32_t r[4];
    rng((uint32_t)blk, (uint32_t)(blk >> 32) ^ (draw << 8), (uint32_t)c, prm.call_lo ^ 0x4B354B35u, r);
    const uint64_t r1 = ((uint64_t)r[1] << 32) | r[0], r2 = ((uint64_t)r[3] << 32) | r[2];
    const double u = ((double)(r1 >> 11) + 0.5) * 0x1.0p-53;  // (0, 1)
    const double gap = floor(log(u) * inv);                   // failures before the next fire
    if (!(gap < (double)kNoiseBlockShots)) break;
    pos += (long long)gap + 1;
    if (pos >= kNoiseBlockShots) break;
    const long long shot = (long long)blk * kNoiseBlockShots + pos;
    if (shot < lo || shot >= hi || (prm.skip_shot0 && shot == 0)) continue;
    // outcome by the conditional distribution: a second uniform on [0, t_last)
    const uint64_t v = __umul64hi(r2, t_last);
    uint32_t k = 0;
    while (k + 1 < m && v >= prm.thresholds[first + k]) ++k;
    const uint64_t* pat = prm.patterns + (size_t)(first + k) * prm.words;
    for (int w = 0; w < prm.words; ++w) {
      const uint64_t pv = pat[w];
      if (pv) atomicXor((unsigned long long*)&prm.f[(shot - lo) * prm.words + w], (unsigned long long)pv);
    }
  }
}

}  // namespace tsb
