// Exact Z[w] * 2^p arithmetic and the float32 tail, device side.
//
// Semantics follow the reference's src/tsim/core/exact_scalar.py:19-137 (int32 wrap-around ring
// product, one power-of-two reduction per combine, reduction to a fixpoint at the end of a fold) and
// the float32 op order fixed in oracle/evaluation.py / oracle/exact_scalar.py.  Every float op uses
// an explicit round-to-nearest intrinsic so that nvcc cannot contract mul+add into FMA.
#pragma once
#include <stdint.h>

namespace tsb {

struct ZW {
  uint32_t c0, c1, c2, c3;  // coefficients of 1, w, i, conj(w); two's complement, wrapping
};

__device__ __forceinline__ ZW zw_make(int a, int b, int c, int d) {
  ZW r;
  r.c0 = (uint32_t)a; r.c1 = (uint32_t)b; r.c2 = (uint32_t)c; r.c3 = (uint32_t)d;
  return r;
}
__device__ __forceinline__ ZW zw_from(int4 v) { return zw_make(v.x, v.y, v.z, v.w); }

// exact_scalar.py:31-39
__device__ __forceinline__ ZW zw_mul(const ZW& x, const ZW& y) {
  ZW r;
  r.c0 = x.c0 * y.c0 + x.c1 * y.c3 - x.c2 * y.c2 + x.c3 * y.c1;
  r.c1 = x.c0 * y.c1 + x.c1 * y.c0 + x.c2 * y.c3 + x.c3 * y.c2;
  r.c2 = x.c0 * y.c2 + x.c1 * y.c1 + x.c2 * y.c0 - x.c3 * y.c3;
  r.c3 = x.c0 * y.c3 - x.c1 * y.c2 - x.c2 * y.c1 + x.c3 * y.c0;
  return r;
}

__device__ __forceinline__ void zw_sar(ZW& c, int sh) {
  c.c0 = (uint32_t)((int32_t)c.c0 >> sh);
  c.c1 = (uint32_t)((int32_t)c.c1 >> sh);
  c.c2 = (uint32_t)((int32_t)c.c2 >> sh);
  c.c3 = (uint32_t)((int32_t)c.c3 >> sh);
}

// exact_scalar.py:42-49: one halving if all coefficients are even and not all zero
__device__ __forceinline__ void zw_reduce1(ZW& c, int& p) {
  uint32_t t = c.c0 | c.c1 | c.c2 | c.c3;
  int red = ((t & 1u) == 0u && t != 0u) ? 1 : 0;
  zw_sar(c, red);
  p += red;
}

// exact_scalar.py:124-137: repeat reduce1 until nothing changes == strip all common factors of two
__device__ __forceinline__ void zw_fixpoint(ZW& c, int& p) {
  uint32_t t = c.c0 | c.c1 | c.c2 | c.c3;
  if (t != 0u) {
    int sh = __ffs((int)t) - 1;
    zw_sar(c, sh);
    p += sh;
  }
}

// int32 (1 << n) with XLA semantics: n >= 32 -> 0
__device__ __forceinline__ uint32_t shl_one(int n) { return n >= 32 ? 0u : (1u << n); }

// exact_scalar.py:74-84
__device__ __forceinline__ void zw_add_p(ZW& c, int& p, const ZW& y, int yp) {
  int d = p - yp;
  uint32_t s1 = d > 0 ? shl_one(d) : 1u;
  uint32_t s2 = d < 0 ? shl_one(-d) : 1u;
  c.c0 = c.c0 * s1 + y.c0 * s2;
  c.c1 = c.c1 * s1 + y.c1 * s2;
  c.c2 = c.c2 * s1 + y.c2 * s2;
  c.c3 = c.c3 * s1 + y.c3 * s2;
  p = min(p, yp);
  zw_reduce1(c, p);
}

// exact float32 2^p: denormals kept, overflow -> +inf, underflow -> 0
__device__ __forceinline__ float pow2_f32(int p) {
  uint32_t bits;
  if (p > 127) bits = 0x7F800000u;
  else if (p >= -126) bits = (uint32_t)(p + 127) << 23;
  else if (p >= -149) bits = 1u << (p + 149);
  else bits = 0u;
  return __uint_as_float(bits);
}

#define TSB_SQRT1_2 __uint_as_float(0x3F3504F3u)

// exact_scalar.py:87-89,218-222 with the op order of oracle/exact_scalar.py:to_complex_parts
__device__ __forceinline__ void zw_to_complex(const ZW& c, int p, float& re, float& im) {
  const float s = TSB_SQRT1_2;
  float f0 = __int2float_rn((int32_t)c.c0), f1 = __int2float_rn((int32_t)c.c1);
  float f2 = __int2float_rn((int32_t)c.c2), f3 = __int2float_rn((int32_t)c.c3);
  float t1 = __fmul_rn(f1, s), t3 = __fmul_rn(f3, s);
  float r = __fadd_rn(__fadd_rn(f0, t1), t3);
  float i = __fsub_rn(__fadd_rn(t1, f2), t3);
  float sc = pow2_f32(p);
  re = __fmul_rn(r, sc);
  im = __fmul_rn(i, sc);
}

// |re + i im| as XLA lowers abs(complex64): max * sqrt(1 + (min/max)^2), min where that is NaN
__device__ __forceinline__ float complex_abs(float re, float im) {
  float a = fabsf(re), b = fabsf(im);
  float mx = fmaxf(a, b), mn = fminf(a, b);
  if (a != a || b != b) { mx = __uint_as_float(0x7FC00000u); mn = mx; }  // NaN in -> NaN out
  float r = __fdiv_rn(mn, mx);
  float t = __fadd_rn(1.0f, __fmul_rn(r, r));
  float res = __fmul_rn(mx, __fsqrt_rn(t));
  return (res != res) ? mn : res;
}

// jax.random threefry2x32 (20 rounds)
__device__ __host__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__device__ __host__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks0 = k0, ks1 = k1, ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += ks0; x1 += ks1;
#define TSB_R(r) x0 += x1; x1 = rotl32(x1, r); x1 ^= x0;
  TSB_R(13) TSB_R(15) TSB_R(26) TSB_R(6)
  x0 += ks1; x1 += ks2 + 1u;
  TSB_R(17) TSB_R(29) TSB_R(16) TSB_R(24)
  x0 += ks2; x1 += ks0 + 2u;
  TSB_R(13) TSB_R(15) TSB_R(26) TSB_R(6)
  x0 += ks0; x1 += ks1 + 3u;
  TSB_R(17) TSB_R(29) TSB_R(16) TSB_R(24)
  x0 += ks1; x1 += ks2 + 4u;
  TSB_R(13) TSB_R(15) TSB_R(26) TSB_R(6)
  x0 += ks2; x1 += ks0 + 5u;
#undef TSB_R
}

// jax.random.uniform(key, float32) element `idx` of a 1-D draw (partitionable threefry)
__device__ __forceinline__ float uniform_f32(uint32_t k0, uint32_t k1, uint64_t idx) {
  uint32_t x0 = (uint32_t)(idx >> 32), x1 = (uint32_t)idx;
  threefry2x32(k0, k1, x0, x1);
  uint32_t bits = x0 ^ x1;
  return __fsub_rn(__uint_as_float((bits >> 9) | 0x3F800000u), 1.0f);
}

}  // namespace tsb
