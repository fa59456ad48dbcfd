/*
 * CPU restatement of tsim's sampling hot path in plain C -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Same algorithm and operation order as the oracle's Python modules (which follows the reference file by file:
 * src/tsim/sampler.py:28-167, compile/evaluate.py:15-59, compile/terms.py:42-207, core/exact_scalar.py:19-137),
 * one shot at a time, reading the MODE_FAITHFUL blob of tsim_b200/pack.py.  It exists so that the CPU baseline of
 * bench.py runs at compiled-code speed (oracle/cport.py spreads shot slices over all host cores with threads -- ctypes
 * releases the GIL); it is validated bit for bit against the NumPy
 * oracle in tests/test_oracle_cport.py.  Nothing in tsim_b200/ links or calls it.
 *
 * float32 tail: one IEEE operation per step, no FMA contraction (-ffp-contract=off), same definitions as
 * oracle/evaluation.py (complex_abs = max*sqrt(1+(min/max)^2), approximate branch summed in graph order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { H_MAGIC, H_VERSION, H_MODE, H_W, H_NUM_F, H_N_OUT, H_N_DIRECT, H_N_COMP, H_N_DRAWS, H_N_LEVELS, H_N_CHUNKS,
       H_MAX_CHUNK, H_OFF_DIRECT, H_OFF_COMP, H_OFF_LEVEL, H_OFF_CHUNK, H_OFF_FSEL, H_OFF_DEST, H_OFF_DATA,
       H_DATA_WORDS, H_TOTAL_WORDS, H_WF64, H_WOUT64 };
enum { COMP_WORDS = 8, LEVEL_WORDS = 12, CHUNK_WORDS = 4, PREFACTOR_WORDS = 8, MAXW = 64 };

typedef struct { uint32_t c[4]; } zw;  /* two's complement, wrapping */

static const int UNIT[8][4] = {{1,0,0,0},{0,1,0,0},{0,0,1,0},{0,0,0,-1},{-1,0,0,0},{0,-1,0,0},{0,0,-1,0},{0,0,0,1}};

static zw zw_mul(zw x, zw y) {
  zw r;
  r.c[0] = x.c[0]*y.c[0] + x.c[1]*y.c[3] - x.c[2]*y.c[2] + x.c[3]*y.c[1];
  r.c[1] = x.c[0]*y.c[1] + x.c[1]*y.c[0] + x.c[2]*y.c[3] + x.c[3]*y.c[2];
  r.c[2] = x.c[0]*y.c[2] + x.c[1]*y.c[1] + x.c[2]*y.c[0] - x.c[3]*y.c[3];
  r.c[3] = x.c[0]*y.c[3] - x.c[1]*y.c[2] - x.c[2]*y.c[1] + x.c[3]*y.c[0];
  return r;
}
static void zw_sar(zw* v, int sh) { for (int i = 0; i < 4; ++i) v->c[i] = (uint32_t)((int32_t)v->c[i] >> sh); }
static void reduce1(zw* v, int32_t* p) {
  uint32_t t = v->c[0] | v->c[1] | v->c[2] | v->c[3];
  if (!(t & 1u) && t) { zw_sar(v, 1); *p += 1; }
}
static void fixpoint(zw* v, int32_t* p) {
  uint32_t t = v->c[0] | v->c[1] | v->c[2] | v->c[3];
  if (t) { int sh = __builtin_ctz(t); zw_sar(v, sh); *p += sh; }
}
static uint32_t shl_one(int64_t n) { return n >= 32 ? 0u : (1u << n); }
static void add_p(zw* c, int32_t* p, zw y, int32_t yp) {
  int64_t d = (int64_t)*p - yp;
  uint32_t s1 = d > 0 ? shl_one(d) : 1u, s2 = d < 0 ? shl_one(-d) : 1u;
  for (int i = 0; i < 4; ++i) c->c[i] = c->c[i]*s1 + y.c[i]*s2;
  if (yp < *p) *p = yp;
  reduce1(c, p);
}
static float pow2f(int64_t p) {
  uint32_t b;
  if (p > 127) b = 0x7F800000u; else if (p >= -126) b = (uint32_t)(p + 127) << 23; else if (p >= -149) b = 1u << (p + 149); else b = 0u;
  float f; memcpy(&f, &b, 4); return f;
}
static void to_complex(zw c, int32_t p, float* re, float* im) {
  uint32_t sb = 0x3F3504F3u; float s; memcpy(&s, &sb, 4);
  volatile float f0 = (float)(int32_t)c.c[0], f1 = (float)(int32_t)c.c[1], f2 = (float)(int32_t)c.c[2], f3 = (float)(int32_t)c.c[3];
  volatile float t1 = f1 * s, t3 = f3 * s;
  volatile float r = f0 + t1; r = r + t3;
  volatile float i = t1 + f2; i = i - t3;
  float sc = pow2f(p);
  *re = r * sc; *im = i * sc;
}
static float cabs_xla(float re, float im) {
  float a = fabsf(re), b = fabsf(im);
  if (a != a || b != b) return NAN;
  float mx = a > b ? a : b, mn = a > b ? b : a;
  volatile float r = mn / mx;
  volatile float t = r * r; t = 1.0f + t;
  volatile float res = mx * sqrtf(t);
  return (res != res) ? mn : res;
}
static uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static void threefry(uint32_t k0, uint32_t k1, uint32_t* x0, uint32_t* x1) {
  static const int RA[4] = {13,15,26,6}, RB[4] = {17,29,16,24};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t a = *x0 + ks[0], b = *x1 + ks[1];
  for (int g = 0; g < 5; ++g) {
    const int* R = (g % 2 == 0) ? RA : RB;
    for (int i = 0; i < 4; ++i) { a += b; b = rotl(b, R[i]); b ^= a; }
    a += ks[(g + 1) % 3]; b += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  *x0 = a; *x1 = b;
}
static float uniform_f32(uint32_t k0, uint32_t k1, uint64_t idx) {
  uint32_t a = (uint32_t)(idx >> 32), b = (uint32_t)idx;
  threefry(k0, k1, &a, &b);
  uint32_t bits = ((a ^ b) >> 9) | 0x3F800000u; float f; memcpy(&f, &bits, 4);
  return f - 1.0f;
}
static int parity(const uint32_t* x, const uint32_t* m, int W) {
  uint32_t t = 0; for (int w = 0; w < W; ++w) t ^= x[w] & m[w];
  return __builtin_popcount(t) & 1;
}

/* One graph's exact product T_g = N * U[phase] * Pp * ff (evaluate.py:40-50) for parameter vector x[W]; advances *rp
 * past the graph's record and hands back the prefactor's power2 and approximate float factor. */
static zw graph_term(const uint32_t** rp, const uint32_t* x, int W, int A, int H, int C, int D, int32_t* Tp, int32_t* power2,
                     float* are, float* aim) {
  const uint32_t* r = *rp;
  zw N = {{1,0,0,0}}; int32_t Np = 0;
  for (int j = 0; j < A; ++j, r += W + 1) {
    int par = parity(x, r, W); uint32_t ctl = r[W];
    zw f = {{1,0,0,0}};
    if (ctl & 8u) { int k = ((par << 2) + (int)ctl) & 7; for (int i = 0; i < 4; ++i) f.c[i] = (uint32_t)(UNIT[k][i] + (i == 0)); }
    if (j == 0) N = f; else { N = zw_mul(N, f); reduce1(&N, &Np); }
  }
  if (A > 0) fixpoint(&N, &Np);
  int h = 0;
  for (int j = 0; j < H; ++j, r += W + 1) h += parity(x, r, W) * (int)(r[W] & 7u);
  int e = 0;
  for (int j = 0; j < C; ++j, r += 2 * W + 1) {
    uint32_t cst = r[2 * W];
    e ^= (parity(x, r, W) ^ (int)(cst & 1u)) & (parity(x, r + W, W) ^ (int)((cst >> 1) & 1u));
  }
  zw Pp = {{1,0,0,0}}; int32_t Pq = 0;
  for (int j = 0; j < D; ++j, r += 2 * W + 1) {
    uint32_t ctl = r[2 * W];
    zw f = {{1,0,0,0}};
    if (ctl & 64u) {
      int a = ((int)(ctl & 7u) + 4 * parity(x, r, W)) & 7, b = ((int)((ctl >> 3) & 7u) + 4 * parity(x, r + W, W)) & 7, cc = (a + b) & 7;
      for (int i = 0; i < 4; ++i) f.c[i] = (uint32_t)((i == 0) + UNIT[a][i] + UNIT[b][i] - UNIT[cc][i]);
    }
    if (j == 0) Pp = f; else { Pp = zw_mul(Pp, f); reduce1(&Pp, &Pq); }
  }
  if (D > 0) fixpoint(&Pp, &Pq);
  int ph = (h + 4 * e + (int)r[0]) & 7;
  zw U, ff; for (int i = 0; i < 4; ++i) { U.c[i] = (uint32_t)UNIT[ph][i]; ff.c[i] = r[1 + i]; }
  *power2 = (int32_t)r[5]; memcpy(are, &r[6], 4); memcpy(aim, &r[7], 4);
  r += PREFACTOR_WORDS;
  *rp = r;
  *Tp = Np + Pq;
  return zw_mul(zw_mul(zw_mul(N, U), Pp), ff);
}

/* |evaluate(level)| for one parameter vector x[W] */
static float eval_level(const uint32_t* blob, const uint32_t* lvl, const uint32_t* x, int W) {
  const uint32_t* chunk_tab = blob + blob[H_OFF_CHUNK];
  const uint32_t* data = blob + blob[H_OFF_DATA];
  const int G = (int)lvl[0], A = (int)lvl[2], H = (int)lvl[3], C = (int)lvl[4], D = (int)lvl[5], approx = lvl[6] & 1;
  if (G == 0) return 0.0f;
  zw S = {{0,0,0,0}}; int32_t Sp = 0; int started = 0; float are_acc = 0.0f, aim_acc = 0.0f;
  for (uint32_t c = lvl[7]; c < lvl[7] + lvl[8]; ++c) {
    const uint32_t* r = data + chunk_tab[c * CHUNK_WORDS];
    int ng = (int)chunk_tab[c * CHUNK_WORDS + 2];
    for (int g = 0; g < ng; ++g) {
      int32_t Tp, power2; float are, aim;
      zw T = graph_term(&r, x, W, A, H, C, D, &Tp, &power2, &are, &aim);
      if (!approx) {
        if (!started) { S = T; Sp = Tp + power2; started = 1; } else add_p(&S, &Sp, T, Tp + power2);
      } else {
        float tre, tim; to_complex(T, Tp, &tre, &tim);
        volatile float a1 = tre * are, a2 = tim * aim, b1 = tre * aim, b2 = tim * are;
        volatile float ure = a1 - a2, uim = b1 + b2;
        float pw = pow2f(power2);
        volatile float vr = ure * pw, vi = uim * pw;
        are_acc = are_acc + vr; aim_acc = aim_acc + vi;
      }
    }
  }
  float re, im;
  if (approx) { re = are_acc; im = aim_acc; } else { fixpoint(&S, &Sp); to_complex(S, Sp, &re, &im); }
  return cabs_xla(re, im);
}

/* ---- float64 "truth" and an alternative float32 lowering, for the margin census (tso_census) -------------------
 * truth: every graph's exact Z[w] value converted to complex double (error ~1e-16 relative), times the stored
 *        complex64 factor widened to double, times 2^power2, summed in double; |.| by hypot.
 * alt:   a different but equally legitimate float32 lowering of evaluate.py:56-59 -- pairwise (tree) reduction over
 *        the graphs, FMA-contracted complex product, |z| = sqrtf(re*re + im*im) -- to count how many draws depend on
 *        the choices XLA is free to make. */
typedef struct { double re, im; } cd;
static cd zw_to_cd(zw c, int32_t p) {
  const double s = 0.70710678118654752440;
  double c0 = (double)(int32_t)c.c[0], c1 = (double)(int32_t)c.c[1], c2 = (double)(int32_t)c.c[2], c3 = (double)(int32_t)c.c[3];
  cd z; z.re = ldexp(c0 + (c1 + c3) * s, p); z.im = ldexp(c2 + (c1 - c3) * s, p);
  return z;
}
#define CENSUS_MAXG 4096
static float tree_sum(const float* v, int n) {  /* pairwise reduction */
  if (n <= 0) return 0.0f;
  if (n == 1) return v[0];
  int m = n / 2;
  volatile float a = tree_sum(v, m), b = tree_sum(v + m, n - m);
  return a + b;
}
static void eval_level_census(const uint32_t* blob, const uint32_t* lvl, const uint32_t* x, int W, double* p64, float* palt) {
  const uint32_t* chunk_tab = blob + blob[H_OFF_CHUNK];
  const uint32_t* data = blob + blob[H_OFF_DATA];
  const int G = (int)lvl[0], A = (int)lvl[2], H = (int)lvl[3], C = (int)lvl[4], D = (int)lvl[5], approx = lvl[6] & 1;
  *p64 = 0.0; *palt = 0.0f;
  if (G == 0) return;
  double sre = 0.0, sim = 0.0;
  static __thread float tr[CENSUS_MAXG], ti[CENSUS_MAXG];
  int n = 0;
  zw S = {{0,0,0,0}}; int32_t Sp = 0; int started = 0;
  for (uint32_t c = lvl[7]; c < lvl[7] + lvl[8]; ++c) {
    const uint32_t* r = data + chunk_tab[c * CHUNK_WORDS];
    int ng = (int)chunk_tab[c * CHUNK_WORDS + 2];
    for (int g = 0; g < ng; ++g) {
      int32_t Tp, power2; float are, aim;
      zw T = graph_term(&r, x, W, A, H, C, D, &Tp, &power2, &are, &aim);
      cd z = zw_to_cd(T, Tp);
      if (approx) {
        double ur = z.re * (double)are - z.im * (double)aim, ui = z.re * (double)aim + z.im * (double)are;
        sre += ldexp(ur, power2); sim += ldexp(ui, power2);
        float tre, tim; to_complex(T, Tp, &tre, &tim);
        float pw = pow2f(power2);
        float ure = fmaf(tre, are, -(tim * aim)), uim = fmaf(tre, aim, tim * are);
        if (n < CENSUS_MAXG) { tr[n] = ure * pw; ti[n] = uim * pw; ++n; }
      } else {
        sre += ldexp(z.re, power2); sim += ldexp(z.im, power2);
        if (!started) { S = T; Sp = Tp + power2; started = 1; } else add_p(&S, &Sp, T, Tp + power2);
      }
    }
  }
  *p64 = hypot(sre, sim);
  float are_, aim_;
  if (approx) {
    are_ = tree_sum(tr, n); aim_ = tree_sum(ti, n);
  } else {
    fixpoint(&S, &Sp); to_complex(S, Sp, &are_, &aim_);
  }
  *palt = sqrtf(fmaf(are_, are_, aim_ * aim_));
}
static void derive_subkeys(uint32_t k0, uint32_t k1, int n, uint32_t* out) {
  for (int j = 0; j < n; ++j) {
    uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 1;
    threefry(k0, k1, &a0, &a1); threefry(k0, k1, &b0, &b1);
    k0 = a0; k1 = a1; out[2*j] = b0; out[2*j+1] = b1;
  }
}

/* sample_program for rows [0, B) of a batch whose first row is in-batch shot `shot_offset`.
 * f: uint8 [B, num_f]; out: uint8 [B, n_out]; norm_dev: float [n_comp] (written when shot_offset == 0).
 * Returns 0, or -1 if the blob is not a MODE_FAITHFUL blob / too wide. */
int tso_sample(const uint32_t* blob, const uint8_t* f, int64_t B, int64_t shot_offset, uint32_t k0, uint32_t k1,
               uint8_t* out, float* norm_dev) {
  if (blob[H_MODE] != 0u || blob[H_W] > MAXW) return -1;
  const int W = (int)blob[H_W], num_f = (int)blob[H_NUM_F], n_out = (int)blob[H_N_OUT];
  const int n_direct = (int)blob[H_N_DIRECT], n_comp = (int)blob[H_N_COMP], n_draws = (int)blob[H_N_DRAWS];
  const uint32_t* direct_tab = blob + blob[H_OFF_DIRECT];
  const uint32_t* comp_tab = blob + blob[H_OFF_COMP];
  const uint32_t* level_tab = blob + blob[H_OFF_LEVEL];
  const uint32_t* fsel = blob + blob[H_OFF_FSEL];
  const uint32_t* dest = blob + blob[H_OFF_DEST];
  uint32_t* subkeys = (uint32_t*)malloc(8 * (size_t)(n_draws > 0 ? n_draws : 1));
  derive_subkeys(k0, k1, n_draws, subkeys);
  for (int64_t s = 0; s < B; ++s) {
    const uint8_t* fr = f + s * num_f;
    uint8_t* o = out + s * n_out;
    memset(o, 0, (size_t)n_out);
    const uint64_t shot = (uint64_t)(shot_offset + s);
    for (int j = 0; j < n_direct; ++j) {
      uint32_t dd = direct_tab[2*j+1];
      o[dd & 0x7FFFFFFFu] = (uint8_t)((fr[direct_tab[2*j]] & 1u) ^ (dd >> 31));
    }
    for (int ci = 0; ci < n_comp; ++ci) {
      const uint32_t* comp = comp_tab + ci * COMP_WORDS;
      const int F = (int)comp[0], n_c = (int)comp[1], first_draw = (int)comp[3];
      const uint32_t* sel = fsel + comp[2];
      uint32_t x[MAXW]; memset(x, 0, sizeof x);
      for (int i = 0; i < F; ++i) x[i >> 5] |= (uint32_t)(fr[sel[i]] & 1u) << (i & 31);
      float prev = 0.0f, dev = 0.0f;
      for (int k = 0; k <= n_c; ++k) {
        const uint32_t* lvl = level_tab + (comp[4] + k) * LEVEL_WORDS;
        const int pos = F + k - 1;
        if (k > 0) x[pos >> 5] |= 1u << (pos & 31);
        float p1 = eval_level(blob, lvl, x, W);
        if (k == 0) { prev = p1; continue; }
        if (shot == 0) {
          uint32_t x0[MAXW]; memcpy(x0, x, sizeof x0); x0[pos >> 5] &= ~(1u << (pos & 31));
          float p0 = eval_level(blob, lvl, x0, W);
          volatile float sum = p0 + p1; volatile float norm = sum / prev; volatile float d = fabsf(norm - 1.0f);
          dev = (dev != dev || d != d) ? NAN : (dev > d ? dev : d);
        }
        float u = uniform_f32(subkeys[2*(first_draw+k-1)], subkeys[2*(first_draw+k-1)+1], shot);
        volatile float q = p1 / prev;
        int bit = u < q;
        volatile float rest = prev - p1;
        prev = bit ? p1 : rest;
        if (!bit) x[pos >> 5] &= ~(1u << (pos & 31)); else o[dest[first_draw + k - 1]] = 1;
      }
      if (shot == 0 && norm_dev) norm_dev[ci] = dev;
    }
  }
  free(subkeys);
  return 0;
}


/* Margin census of the draws of rows [0, B): the float32 path above decides the bits (so prefixes are the reference
 * path's), and every draw is re-judged with the float64 truth and with the alternative float32 lowering.
 *   stats[0] draws            stats[1] draws with |u - q32| <= tol * q32 ("margin draws")
 *   stats[2] draws OUTSIDE the margin whose bit differs under float64 (must be 0 for the margin to mean anything)
 *   stats[3] draws whose bit differs under float64 (all inside the margin if stats[2] == 0)
 *   stats[4] draws whose bit differs under the alternative float32 lowering
 *   stats[5] draws with a non-finite or non-positive ratio (degenerate: prev == 0 etc.), not judged
 *   stats[6], stats[7]: as stats[1], stats[2] for the wider tolerance tol_wide
 * maxrel[0]: max |q32 - q64| / q64 over judged draws; maxrel[1]: the same for the alternative lowering. */
int tso_census(const uint32_t* blob, const uint8_t* f, int64_t B, int64_t shot_offset, uint32_t k0, uint32_t k1, double tol,
               double tol_wide, int64_t* stats, double* maxrel) {
  if (blob[H_MODE] != 0u || blob[H_W] > MAXW) return -1;
  const int W = (int)blob[H_W], num_f = (int)blob[H_NUM_F];
  const int n_comp = (int)blob[H_N_COMP], n_draws = (int)blob[H_N_DRAWS];
  const uint32_t* comp_tab = blob + blob[H_OFF_COMP];
  const uint32_t* level_tab = blob + blob[H_OFF_LEVEL];
  const uint32_t* fsel = blob + blob[H_OFF_FSEL];
  for (uint32_t i = 0; i < blob[H_N_LEVELS]; ++i)
    if (level_tab[i * LEVEL_WORDS] > CENSUS_MAXG) return -2;
  uint32_t* subkeys = (uint32_t*)malloc(8 * (size_t)(n_draws > 0 ? n_draws : 1));
  derive_subkeys(k0, k1, n_draws, subkeys);
  for (int i = 0; i < 8; ++i) stats[i] = 0;
  maxrel[0] = maxrel[1] = 0.0;
  for (int64_t s = 0; s < B; ++s) {
    const uint8_t* fr = f + s * num_f;
    const uint64_t shot = (uint64_t)(shot_offset + s);
    for (int ci = 0; ci < n_comp; ++ci) {
      const uint32_t* comp = comp_tab + ci * COMP_WORDS;
      const int F = (int)comp[0], n_c = (int)comp[1], first_draw = (int)comp[3];
      const uint32_t* sel = fsel + comp[2];
      uint32_t x[MAXW]; memset(x, 0, sizeof x);
      for (int i = 0; i < F; ++i) x[i >> 5] |= (uint32_t)(fr[sel[i]] & 1u) << (i & 31);
      float prev = 0.0f, prev_alt = 0.0f; double prev64 = 0.0;
      for (int k = 0; k <= n_c; ++k) {
        const uint32_t* lvl = level_tab + (comp[4] + k) * LEVEL_WORDS;
        const int pos = F + k - 1;
        if (k > 0) x[pos >> 5] |= 1u << (pos & 31);
        float p1 = eval_level(blob, lvl, x, W), p1_alt; double p1_64;
        eval_level_census(blob, lvl, x, W, &p1_64, &p1_alt);
        if (k == 0) { prev = p1; prev64 = p1_64; prev_alt = p1_alt; continue; }
        float u = uniform_f32(subkeys[2*(first_draw+k-1)], subkeys[2*(first_draw+k-1)+1], shot);
        volatile float q = p1 / prev;
        int bit = u < q;
        double q64 = p1_64 / prev64;
        volatile float qa = p1_alt / prev_alt;
        stats[0] += 1;
        if (!(q64 > 0.0) || !isfinite(q64) || !(q > 0.0f) || !isfinite((double)q)) {
          /* degenerate draw (vanishing marginal on this prefix): nothing to judge unless the two disagree on the bit */
          stats[5] += 1;
          if (bit != ((double)u < q64)) stats[3] += 1;
        } else {
          int in_margin = fabs((double)u - (double)q) <= tol * (double)q;
          int bit64 = (double)u < q64;
          int in_wide = fabs((double)u - (double)q) <= tol_wide * (double)q;
          if (in_margin) stats[1] += 1;
          if (in_wide) stats[6] += 1;
          if (bit64 != bit) { stats[3] += 1; if (!in_margin) stats[2] += 1; if (!in_wide) stats[7] += 1; }
          if ((u < qa) != bit) stats[4] += 1;
          double rel = fabs((double)q - q64) / q64, rela = fabs((double)qa - q64) / q64;
          if (rel > maxrel[0]) maxrel[0] = rel;
          if (rela > maxrel[1]) maxrel[1] = rela;
        }
        volatile float rest = prev - p1; volatile float rest_alt = prev_alt - p1_alt;
        prev = bit ? p1 : rest;
        prev_alt = bit ? p1_alt : rest_alt;
        prev64 = bit ? p1_64 : prev64 - p1_64;
        if (!bit) x[pos >> 5] &= ~(1u << (pos & 31));
      }
    }
  }
  free(subkeys);
  return 0;
}
