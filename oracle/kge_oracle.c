/*
 * oracle/kge_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (pykg2vec_b200/) never does.
 *
 * What it is: an independent, scalar C restatement of the reference's scoring
 * algorithms (Sujit-O/pykg2vec, file:line cited per function, relative to
 * /root/reference/) evaluated in the CANONICAL ARITHMETIC of DESIGN.md §3:
 *
 *   - every operation is IEEE-754 binary32, round-to-nearest-even, no
 *     contraction except where fmaf() is written (build with -ffp-contract=off);
 *   - RSUM: a reduction over the embedding axis keeps 8 partial sums; element j
 *     is accumulated into partial (j>>2)&7 in increasing j; the partials are
 *     combined by the butterfly  P[i] <- P[i] + P[i^4];  then ^2;  then ^1;
 *   - normalisation multiplies by  1.0f / max(sqrt(RSUM(x^2)), 1e-12f);
 *   - sin/cos (RotatE) are the fixed polynomial routines below, not libm;
 *   - GROUPING says which two operands are combined first: TAIL = (h,r) then t,
 *     HEAD = (r,t) then h.  forward() uses TAIL (the reference's own
 *     left-to-right order); the 1-vs-all head sweep uses HEAD.
 *
 * The CUDA kernels implement the same specification with warp shuffles; the two
 * must agree BIT-FOR-BIT on scores, hence exactly on rank counts.
 *
 * Parity pinning: the reference's test-suite holds no golden vector for this
 * path, so this oracle is pinned against outputs of the
 * reference itself, generated in the build container by
 * tests/golden/make_golden.py and committed under tests/golden/*.npz
 * (tests/test_oracle_golden.py: scores within 1e-4 relative, ranks exact).
 * HoLE cannot be executed by the installed torch (legacy torch.ifft removed): its golden
 * vectors come from an emulation of the legacy semantics (tests/golden/make_golden.py) —
 * parity for HoLE is pinned on that emulation, not on a run of the reference's own forward().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KGE_MAX_TABLES 16
enum { KGE_TRANSE = 0, KGE_TRANSH = 1, KGE_TRANSD = 2, KGE_TRANSR = 3, KGE_ROTATE = 4,
       KGE_HOLE = 5, KGE_DISTMULT = 6, KGE_COMPLEX = 7, KGE_CP = 8, KGE_SIMPLE = 9,
       KGE_TRANSM = 10, KGE_RESCAL = 11, KGE_ANALOGY = 12, KGE_SIMPLE_IGNR = 13, KGE_QUATE = 14,
       KGE_OCTONIONE = 15, KGE_KG2E = 16, KGE_SLM = 17, KGE_SME = 18, KGE_SME_BL = 19, KGE_NTN = 20, KGE_CONVKB = 21 };
enum { KGE_GROUP_TAIL = 0, KGE_GROUP_HEAD = 1 };

typedef struct kge_model {
  int32_t model, dim, rel_dim, l1_flag;
  float margin, phase_scale;
  int64_t num_ent, num_rel;
  const float* tables[KGE_MAX_TABLES];
} kge_model_t;

/* ---------------------------------------------------------------- RSUM --- */
typedef struct { float p[8]; } rsum_t;
static inline void rs_init(rsum_t* s) { for (int i = 0; i < 8; ++i) s->p[i] = 0.0f; }
static inline float* rs_at(rsum_t* s, int j) { return &s->p[(j >> 2) & 7]; }
static inline float rs_finish(rsum_t* s) {
  float a[8];
  for (int off = 4; off >= 1; off >>= 1) {
    for (int i = 0; i < 8; ++i) a[i] = s->p[i] + s->p[i ^ off];
    for (int i = 0; i < 8; ++i) s->p[i] = a[i];
  }
  return s->p[0];
}

/* 1 / max(||x||_2, 1e-12)   — F.normalize(x, p=2, dim=-1), eps=1e-12
 * (torch.nn.functional.normalize as called at pairwise.py:69-71) */
static float inv_norm(const float* x, int d) {
  rsum_t s; rs_init(&s);
  for (int j = 0; j < d; ++j) { float* p = rs_at(&s, j); *p = fmaf(x[j], x[j], *p); }
  float n = sqrtf(rs_finish(&s));
  return 1.0f / fmaxf(n, 1e-12f);
}
static float dot(const float* a, const float* b, int d) {
  rsum_t s; rs_init(&s);
  for (int j = 0; j < d; ++j) { float* p = rs_at(&s, j); *p = fmaf(a[j], b[j], *p); }
  return rs_finish(&s);
}

/* ------------------------------------------------ canonical sin / cos ----- */
/* Cody-Waite reduction by pi/2 in three parts + Cephes single-precision
 * minimax polynomials.  Pure fmaf/mul/add: identical on CPU and GPU. */
void kgeo_sincosf(float x, float* sn, float* cs) {
  const float k = rintf(x * 0.636619772367581343f); /* x * 2/pi */
  float r = fmaf(-k, 1.57079601287841796875f, x);
  r = fmaf(-k, 3.13916473303834209219e-7f, r);
  r = fmaf(-k, 5.39030252995776476554e-15f, r);
  const float s = r * r;
  /* sin(r) = r + r*s*(S1 + s*(S2 + s*S3)) */
  float ps = fmaf(s, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(s, ps, -1.6666654611e-1f);
  const float sr = fmaf(r * s, ps, r);
  /* cos(r) = 1 - s/2 + s*s*(C1 + s*(C2 + s*C3)) */
  float pc = fmaf(s, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(s, pc, 4.166664568298827e-2f);
  const float cr = fmaf(s * s, pc, fmaf(s, -0.5f, 1.0f));
  /* k may be large; reduce mod 4 in integer arithmetic (|k| < 2^31 assumed) */
  const int q = ((int)k) & 3;
  float so, co;
  switch (q) {
    case 0: so = sr; co = cr; break;
    case 1: so = cr; co = -sr; break;
    case 2: so = -sr; co = -cr; break;
    default: so = -cr; co = sr; break;
  }
  *sn = so; *cs = co;
}

/* ---------------------------------------------- canonical exp / sigmoid ------ */
/* Cephes expf: k = rint(x*log2e); r = x - k*ln2 (two parts); degree-5 polynomial; scale by 2^k
 * built from the exponent bits.  x is clamped to [-87, 87].  Pure fmaf: identical on the GPU. */
float kgeo_expf(float x) {
  x = fminf(fmaxf(x, -87.0f), 87.0f);
  const float k = rintf(x * 1.44269504088896341f);
  float r = fmaf(-k, 0.693359375f, x);
  r = fmaf(-k, -2.12194440e-4f, r);
  float p = fmaf(r, 1.9875691500e-4f, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  const float y = fmaf(p, r * r, r) + 1.0f;
  union { uint32_t u; float f; } two_k;
  two_k.u = (uint32_t)((int)k + 127) << 23;
  return y * two_k.f;
}
float kgeo_sigmoidf(float x) { return 1.0f / (1.0f + kgeo_expf(-x)); }

/* Canonical tanh (Cephes tanhf: odd polynomial below 0.625, else 1 - 2/(exp(2|x|)+1)) */
float kgeo_tanhf(float x) {
  const float ax = fabsf(x);
  if (ax < 0.625f) {
    const float z = x * x;
    float p = fmaf(-5.70498872745e-3f, z, 2.06390887954e-2f);
    p = fmaf(p, z, -5.37397155531e-2f);
    p = fmaf(p, z, 1.33314422036e-1f);
    p = fmaf(p, z, -3.33332819422e-1f);
    return fmaf(p * z, x, x);
  }
  const float e = kgeo_expf(2.0f * ax);
  const float t = 1.0f - 2.0f * (1.0f / (e + 1.0f));
  return x < 0.0f ? -t : t;
}

/* Canonical natural logarithm (Cephes logf in explicit fma; identical on the GPU).  x <= 0 -> -inf / NaN;
 * subnormal inputs are first scaled by 2^23. */
float kgeo_logf(float x) {
  if (!(x > 0.0f)) return x == 0.0f ? -INFINITY : NAN;
  if (isinf(x)) return x;
  int e = 0;
  if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
  union { float f; uint32_t u; } v; v.f = x;
  e += (int)((v.u >> 23) & 0xff) - 126;
  v.u = (v.u & 0x007fffffu) | 0x3f000000u;  /* mantissa in [0.5, 1) */
  float m = v.f;
  if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
  const float z = m * m;
  float y = fmaf(7.0376836292e-2f, m, -1.1514610310e-1f);
  y = fmaf(y, m, 1.1676998740e-1f);
  y = fmaf(y, m, -1.2420140846e-1f);
  y = fmaf(y, m, 1.4249322787e-1f);
  y = fmaf(y, m, -1.6668057665e-1f);
  y = fmaf(y, m, 2.0000714765e-1f);
  y = fmaf(y, m, -2.4999993993e-1f);
  y = fmaf(y, m, 3.3333331174e-1f);
  y = (y * m) * z;
  const float fe = (float)e;
  y = fmaf(-2.12194440e-4f, fe, y);
  y = fmaf(-0.5f, z, y);
  float r = m + y;
  r = fmaf(0.693359375f, fe, r);
  return r;
}

/* 1 / ||x||_2 without epsilon (KG2E.get_normalized_data pairwise.py:1056-1059, Rescal :862-865) */
static float inv_norm_noeps(const float* x, int d) {
  rsum_t s; rs_init(&s);
  for (int j = 0; j < d; ++j) { float* p = rs_at(&s, j); *p = fmaf(x[j], x[j], *p); }
  return 1.0f / sqrtf(rs_finish(&s));
}

/* Hamilton product, QuatE/OctonionE._qmult (pointwise.py:962-968), canonical fma order */
static void hyper_qmult(const float A[4], const float B[4], float O[4]) {
  O[0] = fmaf(-A[3], B[3], fmaf(-A[2], B[2], fmaf(-A[1], B[1], A[0] * B[0])));
  O[1] = fmaf(-B[2], A[3], fmaf(A[2], B[3], fmaf(B[0], A[1], A[0] * B[1])));
  O[2] = fmaf(-B[3], A[1], fmaf(A[3], B[1], fmaf(B[0], A[2], A[0] * B[2])));
  O[3] = fmaf(-B[1], A[2], fmaf(A[1], B[2], fmaf(B[0], A[3], A[0] * B[3])));
}

/* ------------------------------------------------------- per-model score -- */
static const float* row(const kge_model_t* m, int k, int64_t i, int width) {
  return m->tables[k] + (size_t)i * (size_t)width;
}

/* final translational distance ||a_hat (+) ...||_p given already-transformed
 * h', r', t' (each then L2-normalised) — the shared tail of TransE/H/D/R/M
 * forward(): pairwise.py:69-76, :146-153, :266-273, :463-470, :336-347 */
static float trans_distance(const float* hv, const float* rv, const float* tv, int d, int l1,
                            int grouping) {
  const float ih = inv_norm(hv, d), ir = inv_norm(rv, d), it = inv_norm(tv, d);
  rsum_t s; rs_init(&s);
  for (int j = 0; j < d; ++j) {
    const float hn = hv[j] * ih, rn = rv[j] * ir, tn = tv[j] * it;
    float x;
    if (grouping == KGE_GROUP_TAIL) { const float q = hn + rn; x = q - tn; }
    else { const float q = rn - tn; x = hn + q; }
    float* p = rs_at(&s, j);
    if (l1) *p = *p + fabsf(x); else *p = fmaf(x, x, *p);
  }
  const float acc = rs_finish(&s);
  return l1 ? acc : sqrtf(acc);
}

static float score_one(const kge_model_t* m, int grouping, int64_t h, int64_t r, int64_t t,
                       float* scratch /* >= 4*max(dim,rel_dim) floats */) {
  const int d = m->dim;
  switch (m->model) {
    case KGE_TRANSE: /* pairwise.py:56-93 */
      return trans_distance(row(m, 0, h, d), row(m, 1, r, d), row(m, 0, t, d), d, m->l1_flag, grouping);
    case KGE_TRANSM: { /* pairwise.py:325-347: theta[r] * TransE distance */
      const float dist = trans_distance(row(m, 0, h, d), row(m, 1, r, d), row(m, 0, t, d), d,
                                        m->l1_flag, grouping);
      return m->tables[2][r] * dist;
    }
    case KGE_TRANSH: { /* pairwise.py:143-182: e_perp = e - (e . w~) w~ */
      const float *hv = row(m, 0, h, d), *rv = row(m, 1, r, d), *tv = row(m, 0, t, d),
                  *wv = row(m, 2, r, d);
      float *wn = scratch, *hp = scratch + d, *tp = scratch + 2 * d;
      const float iw = inv_norm(wv, d);
      for (int j = 0; j < d; ++j) wn[j] = wv[j] * iw;
      const float ah = dot(hv, wn, d), at = dot(tv, wn, d);
      for (int j = 0; j < d; ++j) { hp[j] = fmaf(-ah, wn[j], hv[j]); tp[j] = fmaf(-at, wn[j], tv[j]); }
      return trans_distance(hp, rv, tp, d, m->l1_flag, grouping);
    }
    case KGE_TRANSD: { /* pairwise.py:229-278: e' = e + (e . e_m) r_m */
      const float *hv = row(m, 0, h, d), *rv = row(m, 1, r, d), *tv = row(m, 0, t, d);
      const float *hm = row(m, 2, h, d), *tm = row(m, 2, t, d), *rm = row(m, 3, r, d);
      float *hp = scratch, *tp = scratch + d;
      const float ah = dot(hv, hm, d), at = dot(tv, tm, d);
      for (int j = 0; j < d; ++j) { hp[j] = fmaf(ah, rm[j], hv[j]); tp[j] = fmaf(at, rm[j], tv[j]); }
      return trans_distance(hp, rv, tp, d, m->l1_flag, grouping);
    }
    case KGE_TRANSR: { /* pairwise.py:405-470: normalise, project by M_r, normalise again */
      const int dr = m->rel_dim;
      const float *hv = row(m, 0, h, d), *rv = row(m, 1, r, dr), *tv = row(m, 0, t, d);
      const float* M = row(m, 2, r, d * dr); /* view(d_e, d_r): M[j*dr + k] */
      float *hp = scratch, *tp = scratch + dr, *rn = scratch + 2 * dr;
      const float ih = inv_norm(hv, d), it = inv_norm(tv, d), ir = inv_norm(rv, dr);
      for (int k = 0; k < dr; ++k) { hp[k] = 0.0f; tp[k] = 0.0f; rn[k] = rv[k] * ir; }
      for (int j = 0; j < d; ++j) {
        const float hn = hv[j] * ih, tn = tv[j] * it;
        for (int k = 0; k < dr; ++k) {
          hp[k] = fmaf(hn, M[(size_t)j * dr + k], hp[k]);
          tp[k] = fmaf(tn, M[(size_t)j * dr + k], tp[k]);
        }
      }
      return trans_distance(hp, rn, tp, dr, m->l1_flag, grouping);
    }
    case KGE_ROTATE: { /* pairwise.py:765-791 */
      const float *hr = row(m, 0, h, d), *hi = row(m, 1, h, d), *rr = row(m, 2, r, d),
                  *tr = row(m, 0, t, d), *ti = row(m, 1, t, d);
      rsum_t s; rs_init(&s);
      for (int c = 0; c < d; c += 4) { /* chunk-wise: 4 real-part terms, then 4 imaginary-part terms */
        const int n = (d - c) < 4 ? (d - c) : 4;
        float sr[4], si[4];
        for (int e = 0; e < n; ++e) {
          const int j = c + e;
          float im, re; kgeo_sincosf(rr[j] * m->phase_scale, &im, &re);
          if (grouping == KGE_GROUP_TAIL) {   /* |h o r - t|^2 as written */
            const float qr = fmaf(hr[j], re, -(hi[j] * im));
            const float qi = fmaf(hr[j], im, hi[j] * re);
            sr[e] = qr - tr[j]; si[e] = qi - ti[j];
          } else {                            /* |t o conj(r) - h|^2: the same value for the unit rotation */
            const float qr = fmaf(tr[j], re, ti[j] * im);
            const float qi = fmaf(ti[j], re, -(tr[j] * im));
            sr[e] = qr - hr[j]; si[e] = qi - hi[j];
          }
        }
        float* p = rs_at(&s, c);
        for (int e = 0; e < n; ++e) *p = fmaf(sr[e], sr[e], *p);
        for (int e = 0; e < n; ++e) *p = fmaf(si[e], si[e], *p);
      }
      return rs_finish(&s) - m->margin; /* -(margin - sum) */
    }
    case KGE_DISTMULT: /* pointwise.py:444-446 */
    case KGE_CP: {     /* pointwise.py:374-376 (separate subject/object tables) */
      const int cp = (m->model == KGE_CP);
      const float *hv = row(m, 0, h, d), *rv = row(m, 1, r, d), *tv = row(m, cp ? 2 : 0, t, d);
      rsum_t s; rs_init(&s);
      for (int j = 0; j < d; ++j) {
        float* p = rs_at(&s, j);
        if (grouping == KGE_GROUP_TAIL) { const float q = hv[j] * rv[j]; *p = fmaf(q, tv[j], *p); }
        else { const float q = rv[j] * tv[j]; *p = fmaf(hv[j], q, *p); }
      }
      return -rs_finish(&s);
    }
    case KGE_CONVKB: { /* pointwise.py:302-318, collapsed: tables [ent, rel, A(3 x d), c0(1)] (include/kge_b200.h) */
      const float *hv = row(m, 0, h, d), *rv = row(m, 1, r, d), *tv = row(m, 0, t, d);
      const float *ah = m->tables[2], *ar = ah + d, *at = ah + 2 * (size_t)d;
      rsum_t sh, sr, st; rs_init(&sh); rs_init(&sr); rs_init(&st);
      for (int j = 0; j < d; ++j) {
        float* p = rs_at(&sh, j); *p = fmaf(hv[j], ah[j], *p);
        p = rs_at(&sr, j); *p = fmaf(rv[j], ar[j], *p);
        p = rs_at(&st, j); *p = fmaf(tv[j], at[j], *p);
      }
      const float a = rs_finish(&sh), b = rs_finish(&sr), c = rs_finish(&st);
      const float s = (grouping == KGE_GROUP_TAIL) ? ((a + b) + c) : (a + (b + c));
      return s + m->tables[3][0];
    }
    case KGE_COMPLEX: { /* pointwise.py:163-188 */
      const float *hr = row(m, 0, h, d), *hi = row(m, 1, h, d), *rr = row(m, 2, r, d),
                  *ri = row(m, 3, r, d), *tr = row(m, 0, t, d), *ti = row(m, 1, t, d);
      rsum_t s; rs_init(&s);
      for (int c = 0; c < d; c += 4) { /* chunk-wise: 4 real-part terms, then 4 imaginary-part terms */
        const int n = (d - c) < 4 ? (d - c) : 4;
        float qr[4], qi[4];
        for (int e = 0; e < n; ++e) {
          const int j = c + e;
          if (grouping == KGE_GROUP_TAIL) {
            qr[e] = fmaf(hr[j], rr[j], -(hi[j] * ri[j]));
            qi[e] = fmaf(hi[j], rr[j], hr[j] * ri[j]);
          } else {
            qr[e] = fmaf(tr[j], rr[j], ti[j] * ri[j]);
            qi[e] = fmaf(ti[j], rr[j], -(tr[j] * ri[j]));
          }
        }
        float* p = rs_at(&s, c);
        if (grouping == KGE_GROUP_TAIL) {
          for (int e = 0; e < n; ++e) *p = fmaf(qr[e], tr[c + e], *p);
          for (int e = 0; e < n; ++e) *p = fmaf(qi[e], ti[c + e], *p);
        } else {
          for (int e = 0; e < n; ++e) *p = fmaf(hr[c + e], qr[e], *p);
          for (int e = 0; e < n; ++e) *p = fmaf(hi[c + e], qi[e], *p);
        }
      }
      return -rs_finish(&s);
    }
    case KGE_HOLE: {
      /* HoLE.forward pairwise.py:1119-1125 AS WRITTEN for torch<1.7 (torch.conj is a no-op on the
       * real [.,2] view and `*` multiplies (re,im) pairs elementwise), which evaluates
       *   e = circconv(even(h), even(t)),  even(x)[n] = (x[n] + x[(d-n)%d]) / 2,
       *   score = -sigmoid(sum_k r^_k e_k)          
       * re-associated so that the query side is combined first:
       *   TAIL: g[m] = sum_n eh[n] r^[(m+n)%d] (sequential n);  s = RSUM_m g[m] * et[m]
       *   HEAD: g[n] = sum_m et[m] r^[(m+n)%d] (sequential m);  s = RSUM_n eh[n] * g[n] */
      const float *hv = row(m, 0, h, d), *rv = row(m, 1, r, d), *tv = row(m, 0, t, d);
      float *rn = scratch, *eh = scratch + d, *et = scratch + 2 * d, *g = scratch + 3 * d;
      const float ir = inv_norm(rv, d);
      for (int j = 0; j < d; ++j) {
        rn[j] = rv[j] * ir;
        eh[j] = 0.5f * (hv[j] + hv[(d - j) % d]);
        et[j] = 0.5f * (tv[j] + tv[(d - j) % d]);
      }
      const float* qe = (grouping == KGE_GROUP_TAIL) ? eh : et;   /* query-side even part */
      const float* ce = (grouping == KGE_GROUP_TAIL) ? et : eh;   /* candidate-side even part */
      for (int a = 0; a < d; ++a) {
        float acc = 0.0f;
        for (int b = 0; b < d; ++b) acc = fmaf(qe[b], rn[(a + b) % d], acc);
        g[a] = acc;
      }
      rsum_t s; rs_init(&s);
      for (int j = 0; j < d; ++j) {
        float* p = rs_at(&s, j);
        if (grouping == KGE_GROUP_TAIL) *p = fmaf(g[j], ce[j], *p); else *p = fmaf(ce[j], g[j], *p);
      }
      return -kgeo_sigmoidf(rs_finish(&s));
    }
    case KGE_RESCAL: {
      /* Rescal.forward pairwise.py:829-865 on tables that embed() has row-normalised in place
       * (the normalisation is a separate entry point: it mutates the weights, :843-844):
       *   score = - h^T M_r t,   M_r = rel_matrices[r].view(d, d)
       *   TAIL: v_k = sum_j h_j M[j,k] (sequential j);  s = RSUM_k v_k t_k
       *   HEAD: u_j = sum_k M[j,k] t_k (sequential k);  s = RSUM_j h_j u_j */
      const float *hv = row(m, 0, h, d), *tv = row(m, 0, t, d);
      const float* M = row(m, 1, r, d * d);
      float* v = scratch;
      rsum_t s; rs_init(&s);
      if (grouping == KGE_GROUP_TAIL) {
        for (int k = 0; k < d; ++k) v[k] = 0.0f;
        for (int j = 0; j < d; ++j)
          for (int k = 0; k < d; ++k) v[k] = fmaf(hv[j], M[(size_t)j * d + k], v[k]);
        for (int k = 0; k < d; ++k) { float* p = rs_at(&s, k); *p = fmaf(v[k], tv[k], *p); }
      } else {
        for (int j = 0; j < d; ++j) {
          float acc = 0.0f;
          for (int k = 0; k < d; ++k) acc = fmaf(M[(size_t)j * d + k], tv[k], acc);
          v[j] = acc;
        }
        for (int j = 0; j < d; ++j) { float* p = rs_at(&s, j); *p = fmaf(hv[j], v[j], *p); }
      }
      return -rs_finish(&s);
    }
    case KGE_SLM:
    case KGE_NTN: {
      /* SLM.forward/layer pairwise.py:525-541:  -sum_k r^_k tanh((h^ mr1)_k + (t^ mr2)_k)
       * NTN.forward/train_layer pairwise.py:919-960: adds the bilinear tensor term h^T W_k t^ and the bias:
       *   pre_k = ((h^T W_k t^ + (h^ mr1)_k) + (t^ mr2)_k) + br_k
       * tables SLM [ent, rel, mr1(d x k), mr2(d x k)], NTN [.., br(1 x k), mr(k x d*d)]; k = rel_dim.
       * Matrix-vector products accumulate sequentially over the input index; h^T W_k t^ is
       * v_j = sum_i h^_i W_k[i,j] (sequential i), then RSUM_j v_j t^_j.  Grouping-independent. */
      const int K = m->rel_dim;
      const float *hv = row(m, 0, h, d), *rv = row(m, 1, r, K), *tv = row(m, 0, t, d);
      const float *mr1 = m->tables[2], *mr2 = m->tables[3];
      const float ih = inv_norm(hv, d), ir = inv_norm(rv, K), it = inv_norm(tv, d);
      float *hn = scratch, *tn = scratch + d;
      for (int i = 0; i < d; ++i) { hn[i] = hv[i] * ih; tn[i] = tv[i] * it; }
      rsum_t s; rs_init(&s);
      for (int k = 0; k < K; ++k) {
        float a = 0.0f, b = 0.0f;
        for (int i = 0; i < d; ++i) { a = fmaf(hn[i], mr1[(size_t)i * K + k], a); b = fmaf(tn[i], mr2[(size_t)i * K + k], b); }
        float pre;
        if (m->model == KGE_NTN) {
          const float* W = m->tables[5] + (size_t)k * d * d;
          rsum_t q; rs_init(&q);
          for (int j = 0; j < d; ++j) {
            float v = 0.0f;
            for (int i = 0; i < d; ++i) v = fmaf(hn[i], W[(size_t)i * d + j], v);
            float* p = rs_at(&q, j); *p = fmaf(v, tn[j], *p);
          }
          pre = ((rs_finish(&q) + a) + b) + m->tables[4][k];
        } else {
          pre = a + b;
        }
        float* p = rs_at(&s, k);
        *p = fmaf(rv[k] * ir, kgeo_tanhf(pre), *p);
      }
      return -rs_finish(&s);
    }
    case KGE_SME:
    case KGE_SME_BL: {
      /* SME.forward pairwise.py:617-661:  -sum_k (mu1 h^ + mu2 r^ + bu)_k (mv1 t^ + mv2 r^ + bv)_k
       * SME_BL.forward pairwise.py:680-724: +sum_k ((mu1 h^)(mu2 r^) + bu)_k ((mv1 t^)(mv2 r^) + bv)_k
       * tables [ent, rel, mu1, mu2, bu, mv1, mv2, bv], matrices d x d (row k dotted with the vector,
       * sequential over the input index).  Grouping-independent. */
      const float *hv = row(m, 0, h, d), *rv = row(m, 1, r, d), *tv = row(m, 0, t, d);
      const float *mu1 = m->tables[2], *mu2 = m->tables[3], *bu = m->tables[4], *mv1 = m->tables[5],
                  *mv2 = m->tables[6], *bv = m->tables[7];
      const float ih = inv_norm(hv, d), ir = inv_norm(rv, d), it = inv_norm(tv, d);
      float *hn = scratch, *rn = scratch + d, *tn = scratch + 2 * d;
      for (int i = 0; i < d; ++i) { hn[i] = hv[i] * ih; rn[i] = rv[i] * ir; tn[i] = tv[i] * it; }
      rsum_t s; rs_init(&s);
      for (int k = 0; k < d; ++k) {
        float u1 = 0.0f, u2 = 0.0f, v1 = 0.0f, v2 = 0.0f;
        for (int i = 0; i < d; ++i) {
          u1 = fmaf(mu1[(size_t)k * d + i], hn[i], u1); u2 = fmaf(mu2[(size_t)k * d + i], rn[i], u2);
          v1 = fmaf(mv1[(size_t)k * d + i], tn[i], v1); v2 = fmaf(mv2[(size_t)k * d + i], rn[i], v2);
        }
        float gu, gv;
        if (m->model == KGE_SME) { gu = (u1 + u2) + bu[k]; gv = (v1 + v2) + bv[k]; }
        else { gu = u1 * u2 + bu[k]; gv = v1 * v2 + bv[k]; }
        float* p = rs_at(&s, k); *p = fmaf(gu, gv, *p);
      }
      const float tot = rs_finish(&s);
      return m->model == KGE_SME ? -tot : tot;
    }
    case KGE_KG2E: {
      /* KG2E.forward/_cal_score_kl_divergence pairwise.py:1021-1084: all six gathered rows are divided
       * by their L2 norm (no epsilon), then  score = sum (s_h+s_r)/s_t + sum (mu_t-mu_h-mu_r)^2/s_t
       * + sum (log s_t - log(s_h+s_r)) - d.  tables [ent_mu, ent_sigma, rel_mu, rel_sigma].
       * Both groupings use this arithmetic. */
      const float *hm = row(m, 0, h, d), *hs = row(m, 1, h, d), *rm = row(m, 2, r, d), *rsg = row(m, 3, r, d),
                  *tm = row(m, 0, t, d), *ts = row(m, 1, t, d);
      const float ihm = inv_norm_noeps(hm, d), ihs = inv_norm_noeps(hs, d), irm = inv_norm_noeps(rm, d),
                  irs = inv_norm_noeps(rsg, d), itm = inv_norm_noeps(tm, d), its = inv_norm_noeps(ts, d);
      rsum_t T, M, D; rs_init(&T); rs_init(&M); rs_init(&D);
      for (int j = 0; j < d; ++j) {
        const float cs = hs[j] * ihs + rsg[j] * irs;
        const float cm = hm[j] * ihm + rm[j] * irm;
        const float st = ts[j] * its;
        const float x = tm[j] * itm - cm;
        float* p = rs_at(&T, j); *p = *p + cs / st;
        p = rs_at(&M, j); *p = *p + (x * x) / st;
        p = rs_at(&D, j); *p = *p + (kgeo_logf(st) - kgeo_logf(cs));
      }
      return ((rs_finish(&T) + rs_finish(&M)) + rs_finish(&D)) - (float)d;
    }
    case KGE_QUATE:
    case KGE_OCTONIONE: {
      /* QuatE.forward pointwise.py:678-694 / OctonionE.forward pointwise.py:886-899 (+ _qmult, _qstar,
       * _omult, _onorm :962-1002): per embedding dimension the relation (hyper)complex number is
       * normalised to unit modulus, multiplied with the head's, and the result is dotted with the
       * tail's.  tables [ent_1..ent_C, rel_1..rel_C], C = 4 or 8.  Both groupings use this arithmetic. */
      const int C = (m->model == KGE_QUATE) ? 4 : 8;
      rsum_t s; rs_init(&s);
      for (int j = 0; j < d; ++j) {
        float hc[8], rc[8], tc[8], o[8];
        for (int c = 0; c < C; ++c) { hc[c] = row(m, c, h, d)[j]; tc[c] = row(m, c, t, d)[j]; rc[c] = row(m, C + c, r, d)[j]; }
        float den2 = rc[0] * rc[0];
        for (int c = 1; c < C; ++c) den2 = fmaf(rc[c], rc[c], den2);
        const float inv = 1.0f / sqrtf(den2);
        for (int c = 0; c < C; ++c) rc[c] = rc[c] * inv;
        if (C == 4) {
          hyper_qmult(hc, rc, o);
        } else {
          float dstar[4] = {rc[4], -rc[5], -rc[6], -rc[7]}, cstar[4] = {rc[0], -rc[1], -rc[2], -rc[3]};
          float p1[4], p2[4], p3[4], p4[4];
          hyper_qmult(hc, rc, p1);          /* a (x) c   */
          hyper_qmult(dstar, hc + 4, p2);   /* d* (x) b  */
          hyper_qmult(rc + 4, hc, p3);      /* d (x) a   */
          hyper_qmult(hc + 4, cstar, p4);   /* b (x) c*  */
          for (int c = 0; c < 4; ++c) { o[c] = p1[c] - p2[c]; o[4 + c] = p3[c] + p4[c]; }
        }
        float* p = rs_at(&s, j);
        for (int c = 0; c < C; ++c) *p = fmaf(o[c], tc[c], *p);
      }
      return -rs_finish(&s);
    }
    case KGE_ANALOGY: {
      /* ANALOGY.forward pointwise.py:97-104: ComplEx score on the half-width (d/2) tables plus
       * DistMult score on the full-width tables; tables [ent, rel, ent_re, ent_im, rel_re, rel_im] */
      const int d2 = d / 2;
      const float *he = row(m, 0, h, d), *re_ = row(m, 1, r, d), *te = row(m, 0, t, d);
      const float *hr = row(m, 2, h, d2), *hi = row(m, 3, h, d2), *rr = row(m, 4, r, d2), *ri = row(m, 5, r, d2),
                  *tr = row(m, 2, t, d2), *ti = row(m, 3, t, d2);
      rsum_t sc; rs_init(&sc);
      for (int j = 0; j < d2; ++j) {
        float* p = rs_at(&sc, j);
        if (grouping == KGE_GROUP_TAIL) {
          const float qr = fmaf(hr[j], rr[j], -(hi[j] * ri[j]));
          const float qi = fmaf(hi[j], rr[j], hr[j] * ri[j]);
          *p = fmaf(qr, tr[j], *p);
          *p = fmaf(qi, ti[j], *p);
        } else {
          const float qr = fmaf(tr[j], rr[j], ti[j] * ri[j]);
          const float qi = fmaf(ti[j], rr[j], -(tr[j] * ri[j]));
          *p = fmaf(hr[j], qr, *p);
          *p = fmaf(hi[j], qi, *p);
        }
      }
      rsum_t sd; rs_init(&sd);
      for (int j = 0; j < d; ++j) {
        float* p = rs_at(&sd, j);
        if (grouping == KGE_GROUP_TAIL) { const float q = he[j] * re_[j]; *p = fmaf(q, te[j], *p); }
        else { const float q = re_[j] * te[j]; *p = fmaf(he[j], q, *p); }
      }
      return (-rs_finish(&sc)) - rs_finish(&sd);
    }
    case KGE_SIMPLE:
    case KGE_SIMPLE_IGNR: {
      /* SimplE.forward pointwise.py:522-526: init = sum(h1 r1 t1) + sum(h2 r2 t2) / 2 (only the second
       * sum is halved — operator precedence), score = -clamp(init, -20, 20); SimplE_ignr.forward
       * :573-581 has no halving.  tables [ent_head, ent_tail, rel, rel_inv];
       * h1 = ent_head[h], t1 = ent_tail[t], h2 = ent_head[t], t2 = ent_tail[h] (:514-519).
       * One interleaved accumulator; the halving is folded into the query-side factor (exact). */
      const float half = (m->model == KGE_SIMPLE) ? 0.5f : 1.0f;
      const float *h1 = row(m, 0, h, d), *t2 = row(m, 1, h, d), *h2 = row(m, 0, t, d), *t1 = row(m, 1, t, d);
      const float *r1 = row(m, 2, r, d), *r2 = row(m, 3, r, d);
      rsum_t s; rs_init(&s);
      for (int c = 0; c < d; c += 4) { /* chunk-wise: 4 terms of the first product, then 4 of the second */
        const int n = (d - c) < 4 ? (d - c) : 4;
        float q1[4], q2[4];
        for (int e = 0; e < n; ++e) {
          const int j = c + e;
          if (grouping == KGE_GROUP_TAIL) {  /* query (h, r): q1 = h1 r1, q2 = half t2 r2 */
            q1[e] = h1[j] * r1[j];
            q2[e] = (t2[j] * r2[j]) * half;
          } else {                           /* query (r, t): q1 = r1 t1, q2 = half r2 h2 */
            q1[e] = r1[j] * t1[j];
            q2[e] = (r2[j] * h2[j]) * half;
          }
        }
        float* p = rs_at(&s, c);
        if (grouping == KGE_GROUP_TAIL) {
          for (int e = 0; e < n; ++e) *p = fmaf(q1[e], t1[c + e], *p);
          for (int e = 0; e < n; ++e) *p = fmaf(q2[e], h2[c + e], *p);
        } else {
          for (int e = 0; e < n; ++e) *p = fmaf(h1[c + e], q1[e], *p);
          for (int e = 0; e < n; ++e) *p = fmaf(t2[c + e], q2[e], *p);
        }
      }
      const float init = rs_finish(&s);
      return -fminf(fmaxf(init, -20.0f), 20.0f);
    }
    default: return NAN;
  }
}

/* Rescal.get_normalized_data pairwise.py:862-865: every row divided by its L2 norm (no epsilon),
 * applied in place to a [rows, width] table. */
int kgeo_normalize_rows(float* table, int64_t rows, int64_t width) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < rows; ++i) {
    float* x = table + i * width;
    rsum_t s; rs_init(&s);
    for (int64_t j = 0; j < width; ++j) { float* p = rs_at(&s, (int)j); *p = fmaf(x[j], x[j], *p); }
    const float inv = 1.0f / sqrtf(rs_finish(&s));
    for (int64_t j = 0; j < width; ++j) x[j] = x[j] * inv;
  }
  return 0;
}

static int scratch_floats(const kge_model_t* m) {
  int w = m->dim > m->rel_dim ? m->dim : m->rel_dim;
  return 4 * w + 16;
}

/* model.forward(h, r, t) for a batch — see kge_score_fwd in include/kge_b200.h */
int kgeo_score_fwd(const kge_model_t* m, int grouping, const int64_t* h, const int64_t* r,
                   const int64_t* t, int64_t n, float* scores) {
  if (!m || !h || !r || !t || !scores) return -1;
  const int sf = scratch_floats(m);
#pragma omp parallel
  {
    float* scratch = (float*)malloc(sizeof(float) * (size_t)sf);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; ++i) scores[i] = score_one(m, grouping, h[i], r[i], t[i], scratch);
    free(scratch);
  }
  return 0;
}

/* ---------------------------------------------------------------- losses -- */
/* Criterion.pairwise_hinge, criterion.py:26-29.  terms[i] (optional) receives the
 * per-pair value, which the kernels must match bit-for-bit; the total is
 * accumulated in double. */
int kgeo_loss_pairwise_hinge(const float* pos, const float* neg, int64_t n, float margin,
                             float* loss_out, float* terms) {
  double acc = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const float a = pos[i] + margin;
    const float v = fmaxf(a - neg[i], 0.0f);
    if (terms) terms[i] = v;
    acc += (double)v;
  }
  *loss_out = (float)acc;
  return 0;
}
static double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); } /* F.softplus threshold=20 */
static double logsigmoid_d(double x) { return x < 0 ? x - log1p(exp(x)) : -log1p(exp(-x)); }
/* Criterion.pointwise_logistic, criterion.py:32-34 */
int kgeo_loss_pointwise_logistic(const float* preds, const float* target, int64_t n, float* loss_out) {
  double acc = 0.0;
  for (int64_t i = 0; i < n; ++i) acc += softplus_d((double)(target[i] * preds[i]));
  *loss_out = (float)(acc / (double)n);
  return 0;
}
/* Criterion.pariwise_logistic, criterion.py:14-23 */
int kgeo_loss_selfadv(const float* pos, const float* neg, int64_t B, int32_t neg_rate, float alpha,
                      float* loss_out) {
  double accp = 0.0, accn = 0.0;
  for (int64_t i = 0; i < B; ++i) {
    accp += logsigmoid_d(-(double)pos[i]);
    const float* ng = neg + i * neg_rate;
    double mx = -INFINITY;
    for (int j = 0; j < neg_rate; ++j) { double z = -(double)ng[j] * alpha; if (z > mx) mx = z; }
    double den = 0.0;
    for (int j = 0; j < neg_rate; ++j) den += exp(-(double)ng[j] * alpha - mx);
    double row = 0.0;
    for (int j = 0; j < neg_rate; ++j) {
      const double w = exp(-(double)ng[j] * alpha - mx) / den;
      row += w * logsigmoid_d((double)ng[j]); /* logsigmoid(-(-neg)) */
    }
    accn += row;
  }
  *loss_out = (float)(-(accn / (double)B) - (accp / (double)B));
  return 0;
}

/* ----------------------------------------------------------- 1-vs-all ----- */
/* Evaluator.test (evaluator.py:309-334) + MetricCalculator.get_tail_rank /
 * get_head_rank (evaluator.py:70-123), in the count formulation that is equal
 * to the reference's sorted-list walk whenever scores are tie-free:
 *   rank0 = #{e : s_e < s_target};   filtered = rank0 - #{e in filter, e != target : s_e < s_target}
 * Candidates are global entity rows [row_lo,row_hi); tables in `m` hold ALL rows
 * (the oracle is never sharded; row_lo/row_hi let tests check partial counts).
 * counts[Q][4] = (tail raw, tail filtered, head raw, head filtered), accumulated. */
int kgeo_rank_1vsall(const kge_model_t* m, int64_t row_lo, int64_t row_hi, const int64_t* qh,
                     const int64_t* qr, const int64_t* qt, int64_t Q, const int64_t* filt_t_ptr,
                     const int64_t* filt_t_idx, const int64_t* filt_h_ptr, const int64_t* filt_h_idx,
                     int32_t* counts) {
  const int sf = scratch_floats(m);
#pragma omp parallel
  {
    float* scratch = (float*)malloc(sizeof(float) * (size_t)sf);
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < Q; ++i) {
      const int64_t h = qh[i], r = qr[i], t = qt[i];
      /* tail direction: candidates replace t, grouping TAIL */
      const float thr_t = score_one(m, KGE_GROUP_TAIL, h, r, t, scratch);
      int32_t raw = 0;
      for (int64_t e = row_lo; e < row_hi; ++e)
        raw += score_one(m, KGE_GROUP_TAIL, h, r, e, scratch) < thr_t;
      int32_t sub = 0;
      if (filt_t_ptr)
        for (int64_t k = filt_t_ptr[i]; k < filt_t_ptr[i + 1]; ++k) {
          const int64_t e = filt_t_idx[k];
          if (e == t || e < row_lo || e >= row_hi) continue;
          sub += score_one(m, KGE_GROUP_TAIL, h, r, e, scratch) < thr_t;
        }
      counts[i * 4 + 0] += raw;
      counts[i * 4 + 1] += raw - sub;
      /* head direction: candidates replace h, grouping HEAD */
      const float thr_h = score_one(m, KGE_GROUP_HEAD, h, r, t, scratch);
      raw = 0;
      for (int64_t e = row_lo; e < row_hi; ++e)
        raw += score_one(m, KGE_GROUP_HEAD, e, r, t, scratch) < thr_h;
      sub = 0;
      if (filt_h_ptr)
        for (int64_t k = filt_h_ptr[i]; k < filt_h_ptr[i + 1]; ++k) {
          const int64_t e = filt_h_idx[k];
          if (e == h || e < row_lo || e >= row_hi) continue;
          sub += score_one(m, KGE_GROUP_HEAD, e, r, t, scratch) < thr_h;
        }
      counts[i * 4 + 2] += raw;
      counts[i * 4 + 3] += raw - sub;
    }
    free(scratch);
  }
  return 0;
}

/* all N scores of one query direction (used to cross-check the count formulation
 * against the reference's topk + list walk on golden data) */
int kgeo_sweep_scores(const kge_model_t* m, int grouping, int64_t h, int64_t r, int64_t t,
                      int64_t row_lo, int64_t row_hi, float* out) {
  const int sf = scratch_floats(m);
#pragma omp parallel
  {
    float* scratch = (float*)malloc(sizeof(float) * (size_t)sf);
#pragma omp for schedule(static)
    for (int64_t e = row_lo; e < row_hi; ++e)
      out[e - row_lo] = (grouping == KGE_GROUP_TAIL) ? score_one(m, grouping, h, r, e, scratch)
                                                     : score_one(m, grouping, e, r, t, scratch);
    free(scratch);
  }
  return 0;
}

/* ------------------------------------------------------ negative sampling -- */
/* process_function_pairwise / _pointwise (generator.py:42-158) with the counter-based generator
 * specified in include/kge_b200.h (kge_sample_negatives).  The positive set is a sorted array of
 * packed keys searched by bisection (independent of the device's hash table). */
static uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static uint64_t pack_key(int64_t h, int64_t r, int64_t t) {
  return ((uint64_t)h << 42) | ((uint64_t)r << 22) | (uint64_t)t;
}
static int cmp_u64(const void* a, const void* b) {
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}
static int contains(const uint64_t* keys, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
  return lo < n && keys[lo] == key;
}
int kgeo_sample_negatives(const int64_t* th, const int64_t* tr, const int64_t* tt, int64_t ntrain,
                          const int64_t* ph, const int64_t* pr, const int64_t* pt, int64_t B,
                          int32_t neg_rate, const float* head_prob, int64_t num_ent, uint64_t seed,
                          uint64_t step, int32_t layout, int64_t* oh, int64_t* orr, int64_t* ot, int64_t* oy) {
  uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(ntrain > 0 ? ntrain : 1));
  for (int64_t i = 0; i < ntrain; ++i) keys[i] = pack_key(th[i], tr[i], tt[i]);
  qsort(keys, (size_t)ntrain, sizeof(uint64_t), cmp_u64);
  const uint64_t base = mix64(seed ^ mix64(step));
  for (int64_t idx = 0; idx < B * neg_rate; ++idx) {
    const int64_t i = idx / neg_rate, j = idx % neg_rate;
    const int64_t h = ph[i], r = pr[i], t = pt[i];
    const uint64_t s0 = mix64(base + (uint64_t)idx);
    const float u = (float)(s0 >> 40) * (1.0f / 16777216.0f);
    const float prob = head_prob ? head_prob[r] : 0.5f;
    const int corrupt_tail = u > prob;
    int64_t e = 0;
    for (int a = 0; a < 64; ++a) {
      e = (int64_t)(((unsigned __int128)mix64(s0 + (uint64_t)a + 1ull) * (unsigned __int128)(uint64_t)num_ent) >> 64);
      const uint64_t key = corrupt_tail ? pack_key(h, r, e) : pack_key(e, r, t);
      if (!contains(keys, ntrain, key)) break;
    }
    const int64_t o = layout == 0 ? idx : i * (1 + neg_rate) + 1 + j;
    oh[o] = corrupt_tail ? h : e; orr[o] = r; ot[o] = corrupt_tail ? e : t;
    if (layout == 1) {
      oy[o] = -1;
      if (j == 0) { const int64_t p = i * (1 + neg_rate); oh[p] = h; orr[p] = r; ot[p] = t; oy[p] = 1; }
    }
  }
  free(keys);
  return 0;
}

/* ------------------------------------------------ projection-model tail ---- */
/* The last layer every projection model of the reference shares:
 *     preds = sigmoid(x . E^T + b)          ConvE.inner_forward projection.py:100-102
 * (same shape in TuckER :335-336, InteractE :444-447, HypER :607-609, AcrE :735-738 and, without
 * bias, ProjE_pointwise.g :248-256).  x is [B,k] (the trunk's output), E the [N,k] entity table,
 * b a [N] bias row (NULL = none).
 * Canonical arithmetic of this path (rule 8): the dot product is ONE sequential
 * fmaf chain over j = 0..k-1 starting from 0.0f (the order in which a register-tiled GEMM
 * accumulates an output element), then one add of the bias, then the canonical sigmoid. */
static float proj_pred(const float* x, const float* e, int k, const float* bias, int64_t n) {
  float acc = 0.0f;
  for (int j = 0; j < k; ++j) acc = fmaf(x[j], e[j], acc);
  if (bias) acc = acc + bias[n];
  return kgeo_sigmoidf(acc);
}

int kgeo_proj_tail_fwd(const float* x, const float* ent, const float* bias, int64_t B, int64_t N,
                       int32_t k, float* preds) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t n = 0; n < N; ++n)
      preds[b * N + n] = proj_pred(x + b * k, ent + n * k, k, bias, n);
  return 0;
}

/* One direction of Criterion.multi_class_bce (criterion.py:41-50):
 *     y    = labels * label_scale + label_shift        (:43-45; scale = 1 - smoothing, shift = 1/N)
 *     loss = mean_{b,n} BCEWithLogits(preds, y)         (:46-47 — applied to the ALREADY sigmoided
 *                                                       preds, as the reference does)
 * BCEWithLogits(z, y) = (1 - y) z + softplus(-z).  grad_preds = grad_scale * d loss / d preds
 *                     = grad_scale * (sigmoid(z) - y) / (B N).
 * The per-element terms use the canonical exp / log / sigmoid; the sum is taken in double here
 * (the CUDA kernel sums fp32 partials: compared with a tolerance, not bit for bit). */
int kgeo_proj_bce(const float* preds, const float* labels, int64_t B, int64_t N, float label_scale,
                  float label_shift, float grad_scale, float* loss_out, float* grad_preds) {
  const int64_t n = B * N;
  const float gs = (float)((double)grad_scale / ((double)B * (double)N));
  double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
  for (int64_t i = 0; i < n; ++i) {
    const float z = preds[i];
    const float y = labels[i] * label_scale + label_shift;
    const float sp = fmaxf(-z, 0.0f) + kgeo_logf(1.0f + kgeo_expf(-fabsf(z)));
    acc += (double)fmaf(1.0f - y, z, sp);
    if (grad_preds) grad_preds[i] = (kgeo_sigmoidf(z) - y) * gs;
  }
  loss_out[0] = (float)(acc / (double)n);
  return 0;
}

/* Backward of the tail: with g = grad_preds * preds * (1 - preds)  (d sigmoid),
 *     grad_x[b,:] += sum_n g[b,n] E[n,:];  grad_ent[n,:] += sum_b g[b,n] x[b,:];  grad_bias[n] += sum_b g[b,n]
 * accumulated in double (a tolerance reference for the CUDA GEMMs, whose summation order is free). */
int kgeo_proj_tail_bwd(const float* grad_preds, const float* preds, const float* x, const float* ent,
                       int64_t B, int64_t N, int32_t k, float* grad_x, float* grad_ent, float* grad_bias) {
  double* gx = (double*)calloc((size_t)(B * k), sizeof(double));
  double* ge = (double*)calloc((size_t)(N * k), sizeof(double));
  double* gb = (double*)calloc((size_t)N, sizeof(double));
  for (int64_t b = 0; b < B; ++b)
    for (int64_t n = 0; n < N; ++n) {
      const float p = preds[b * N + n];
      const double g = (double)(grad_preds[b * N + n] * (p * (1.0f - p)));
      gb[n] += g;
      for (int j = 0; j < k; ++j) {
        gx[b * k + j] += g * (double)ent[n * k + j];
        ge[n * k + j] += g * (double)x[b * k + j];
      }
    }
  if (grad_x) for (int64_t i = 0; i < B * k; ++i) grad_x[i] += (float)gx[i];
  if (grad_ent) for (int64_t i = 0; i < N * k; ++i) grad_ent[i] += (float)ge[i];
  if (grad_bias) for (int64_t i = 0; i < N; ++i) grad_bias[i] += (float)gb[i];
  free(gx); free(ge); free(gb);
  return 0;
}

/* predict_tail_rank / predict_head_rank (projection.py:119-125: topk of -preds over all N) walked
 * by MetricCalculator.get_*_rank (evaluator.py:70-123), in the count formulation:
 *   rank0 = #{n : preds[q,n] > preds[q,tgt]};  filtered = rank0 - #{n in filter, n != tgt : same}
 * counts[q*4 + 2*dir + {0,1}] accumulated (dir 0 = tail -> columns 0,1; 1 = head -> 2,3). */
int kgeo_proj_rank(const float* x, const float* ent, const float* bias, int64_t Q, int64_t N, int32_t k,
                   const int64_t* tgt, const int64_t* filt_ptr, const int64_t* filt_idx, int32_t dir,
                   int32_t* counts) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t q = 0; q < Q; ++q) {
    const float* xq = x + q * k;
    const int64_t t = tgt[q];
    const float thr = proj_pred(xq, ent + t * k, k, bias, t);
    int32_t raw = 0, sub = 0;
    for (int64_t n = 0; n < N; ++n) raw += proj_pred(xq, ent + n * k, k, bias, n) > thr;
    if (filt_ptr)
      for (int64_t p = filt_ptr[q]; p < filt_ptr[q + 1]; ++p) {
        const int64_t e = filt_idx[p];
        if (e == t) continue;
        sub += proj_pred(xq, ent + e * k, k, bias, e) > thr;
      }
    counts[q * 4 + 2 * dir + 0] += raw;
    counts[q * 4 + 2 * dir + 1] += raw - sub;
  }
  return 0;
}


/* ConvE trunk in inference mode: ConvE.forward + inner_forward up to the x.E^T product
 * (projection.py:104-112, :86-99 with self.training == False).  Canonical arithmetic as
 * pykg2vec_b200/csrc/kge_conve.cuh states it (BatchNorm folded to fma(v, a, c); conv = 9
 * sequential fmaf from 0 then + bias; Linear = one sequential fmaf chain per slice of 512
 * consecutive features, the slice sums added in ascending order, then + bias). */
typedef struct kge_conve {
  int32_t hidden_size, hidden_size_1;
  float bn0_eps, bn1_eps;
  const float* ent; const float* rel;
  const float* bn0_weight; const float* bn0_bias; const float* bn0_mean; const float* bn0_var;
  const float* conv_weight; const float* conv_bias;
  const float* bn1_weight; const float* bn1_bias; const float* bn1_mean; const float* bn1_var;
  const float* fc_weight; const float* fc_bias;
} kge_conve_t;

static void bn_fold(float w, float b, float mean, float var, float eps, float* a, float* c) {
  *a = w / sqrtf(var + eps);
  *c = b - mean * *a;
}

int kgeo_conve_trunk_fwd(const kge_conve_t* p, const int64_t* e, const int64_t* r, int64_t Q, float* x) {
  const int k = p->hidden_size, W = p->hidden_size_1, h2 = k / W, H = 2 * h2;
  const int Ho = H - 2, Wo = W - 2, plane = Ho * Wo;
  const int64_t F = (int64_t)32 * plane;
  float a0, c0, a1[32], c1[32];
  bn_fold(p->bn0_weight[0], p->bn0_bias[0], p->bn0_mean[0], p->bn0_var[0], p->bn0_eps, &a0, &c0);
  for (int c = 0; c < 32; ++c)
    bn_fold(p->bn1_weight[c], p->bn1_bias[c], p->bn1_mean[c], p->bn1_var[c], p->bn1_eps, &a1[c], &c1[c]);
#pragma omp parallel
  {
    float* img = (float*)malloc(sizeof(float) * (size_t)(H * W));
    float* feat = (float*)malloc(sizeof(float) * (size_t)F);
#pragma omp for schedule(static)
    for (int64_t q = 0; q < Q; ++q) {
      const float* er = p->ent + e[q] * k;
      const float* rr = p->rel + r[q] * k;
      for (int i = 0; i < h2 * W; ++i) {   /* torch.cat([e.view(h2,W), r.view(h2,W)], dim=height) */
        img[i] = fmaf(er[i], a0, c0);
        img[h2 * W + i] = fmaf(rr[i], a0, c0);
      }
      for (int c = 0; c < 32; ++c)
        for (int i = 0; i < Ho; ++i)
          for (int j = 0; j < Wo; ++j) {
            float acc = 0.0f;
            for (int di = 0; di < 3; ++di)
              for (int dj = 0; dj < 3; ++dj)
                acc = fmaf(p->conv_weight[c * 9 + di * 3 + dj], img[(i + di) * W + j + dj], acc);
            feat[c * plane + i * Wo + j] = fmaxf(fmaf(acc + p->conv_bias[c], a1[c], c1[c]), 0.0f);
          }
      for (int n = 0; n < k; ++n) {   /* Linear: per-slice fmaf chains, slice sums added in order */
        const float* w = p->fc_weight + (int64_t)n * F;
        float total = 0.0f;
        for (int64_t f0 = 0; f0 < F; f0 += 512) {
          const int64_t f1 = f0 + 512 < F ? f0 + 512 : F;
          float acc = 0.0f;
          for (int64_t f = f0; f < f1; ++f) acc = fmaf(feat[f], w[f], acc);
          total = (f0 == 0) ? acc : total + acc;
        }
        x[q * k + n] = fmaxf(total + p->fc_bias[n], 0.0f);
      }
    }
    free(img); free(feat);
  }
  return 0;
}

int kgeo_abi_version(void) { return 3; }
