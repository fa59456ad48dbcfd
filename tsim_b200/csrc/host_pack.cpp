// Host-side bit packing of the reference's f matrix (uint8[B, num_f], values 0/1) into the device's 64-bit row words,
// SIMD variants with run-time dispatch.  Plain C++ (g++), linked into libtsim_b200.so; called from the host pipeline of
// tsb_sample_host when the caller hands over byte rows (tsim's ChannelSampler.sample format, noise/channels.py:624-658).
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

namespace {

// eight bytes -> eight bits with one multiply: byte k sits at bit 8k, the magic constant moves its LSB to bit 56 + k
inline uint64_t pack_word_scalar(const uint8_t* p, int cols) {
  uint64_t v = 0;
  int b = 0;
  for (; b + 8 <= cols; b += 8) {
    uint64_t x;
    memcpy(&x, p + b, 8);
    v |= (((x & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56) << b;
  }
  for (; b < cols; ++b) v |= (uint64_t)(p[b] & 1u) << b;
  return v;
}

void pack_scalar(const uint8_t* src, long long n, int num_f, int wf, uint64_t* dst) {
  for (long long r = 0; r < n; ++r)
    for (int w = 0; w < wf; ++w) dst[r * wf + w] = pack_word_scalar(src + r * num_f + 64 * w, std::min(64, num_f - 64 * w));
}

// `avail`: bytes readable from src on (the rows that follow belong to the same array), so that full-width loads of a
// ragged row stay inside the caller's buffer; the last rows fall back to the scalar form.
__attribute__((target("avx2"))) void pack_avx2(const uint8_t* src, long long n, int num_f, int wf, uint64_t* dst, long long avail) {
  for (long long r = 0; r < n; ++r) {
    for (int w = 0; w < wf; ++w) {
      const long long off = r * num_f + 64 * w;
      const int cols = std::min(64, num_f - 64 * w);
      if (off + 64 <= avail) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + off));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + off + 32));
        const uint64_t lo = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(a, 7));
        const uint64_t hi = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(b, 7));
        const uint64_t m = cols == 64 ? ~0ull : ((1ull << cols) - 1ull);
        dst[r * wf + w] = (lo | (hi << 32)) & m;
      } else {
        dst[r * wf + w] = pack_word_scalar(src + off, cols);
      }
    }
  }
}

__attribute__((target("avx512f,avx512bw"))) void pack_avx512(const uint8_t* src, long long n, int num_f, int wf, uint64_t* dst, long long avail) {
  const __m512i one = _mm512_set1_epi8(1);
  for (long long r = 0; r < n; ++r) {
    for (int w = 0; w < wf; ++w) {
      const long long off = r * num_f + 64 * w;
      const int cols = std::min(64, num_f - 64 * w);
      if (off + 64 <= avail) {
        const __m512i v = _mm512_loadu_si512(src + off);
        const uint64_t m = cols == 64 ? ~0ull : ((1ull << cols) - 1ull);
        dst[r * wf + w] = (uint64_t)_mm512_test_epi8_mask(v, one) & m;
      } else {
        dst[r * wf + w] = pack_word_scalar(src + off, cols);
      }
    }
  }
}

}  // namespace

extern "C" {

// 0 scalar, 1 avx2, 2 avx512bw: what this host runs
int tsb_host_pack_isa(void) {
  static const int isa = [] {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f")) return 2;
    if (__builtin_cpu_supports("avx2")) return 1;
    return 0;
  }();
  return isa;
}

void tsb_host_pack_rows(const uint8_t* src, long long n, int num_f, int wf, uint64_t* dst, long long avail, int isa) {
  if (isa < 0) isa = tsb_host_pack_isa();
  isa = std::min(isa, tsb_host_pack_isa());
  if (isa >= 2) pack_avx512(src, n, num_f, wf, dst, avail);
  else if (isa == 1) pack_avx2(src, n, num_f, wf, dst, avail);
  else pack_scalar(src, n, num_f, wf, dst);
}

}  // extern "C"
