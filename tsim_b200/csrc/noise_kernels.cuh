// K5: error-mechanism (channel) sampler on the device.
//
// Replaces the host loop of the reference's ChannelSampler.sample (src/tsim/noise/channels.py:624-658):
// every channel fires in a shot with probability p_fire; a fired channel picks a non-identity outcome
// from its conditional distribution and XORs that outcome's precomputed f-pattern into the shot's row.
// The reference draws geometric skips from NumPy's PCG64 stream, which cannot be reproduced in parallel,
// so parity here is statistical (same distribution, different stream): each (shot, channel) pair gets
// its own 64-bit uniform from a counter-based Philox4x32-10 generator keyed by (seed, call number), which
// makes the f rows independent of batch partitioning and of the number of GPUs.
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

struct NoiseParams {
  const uint32_t* __restrict__ chan;        // [n_channels][2]: first outcome index, number of non-identity outcomes
  const uint64_t* __restrict__ thresholds;  // [n_outcomes_total] cumulative: fire iff r < thresholds[last of channel]
  const uint64_t* __restrict__ patterns;    // [n_outcomes_total][words]
  uint64_t* __restrict__ f;                 // [B][words], zero-initialised by the caller
  long long B;
  long long shot_offset;
  int n_channels;
  int words;
  int chan_per_thread;  // channels handled by one thread (even)
  uint32_t seed_lo, seed_hi;
  uint32_t call_lo, call_hi;
  int skip_shot0;       // leave in-batch shot 0 noiseless (reference-sample row, sampler.py:395-396)
};

// grid: x over shots, y over channel groups
__global__ void __launch_bounds__(256) noise_kernel(const NoiseParams prm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= prm.B) return;
  const unsigned long long shot = (unsigned long long)(prm.shot_offset + i);
  if (prm.skip_shot0 && shot == 0ull) return;
  const int c_lo = blockIdx.y * prm.chan_per_thread;
  const int c_hi = min(prm.n_channels, c_lo + prm.chan_per_thread);
  const Philox rng{prm.seed_lo ^ prm.call_lo * 0x9E3779B9u, prm.seed_hi ^ prm.call_hi};
  for (int c = c_lo; c < c_hi; c += 2) {
    uint32_t r[4];
    rng((uint32_t)shot, (uint32_t)(shot >> 32), (uint32_t)(c >> 1), prm.call_lo, r);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cc = c + h;
      if (cc >= c_hi) break;
      const uint64_t u = ((uint64_t)r[2 * h + 1] << 32) | r[2 * h];
      const uint32_t first = prm.chan[2 * cc], m = prm.chan[2 * cc + 1];
      if (u < prm.thresholds[first + m - 1]) {  // fired (rare)
        uint32_t k = 0;
        while (k + 1 < m && u >= prm.thresholds[first + k]) ++k;
        const uint64_t* pat = prm.patterns + (size_t)(first + k) * prm.words;
        for (int w = 0; w < prm.words; ++w) {
          const uint64_t v = pat[w];
          if (v) atomicXor((unsigned long long*)&prm.f[i * prm.words + w], (unsigned long long)v);
        }
      }
    }
  }
}

}  // namespace tsb
