// kge_project.cuh — candidate-side projection of EVERY entity row for one relation.
//
// TransH and TransD project the entity rows with a vector that depends on the relation before
// the TransE distance (pykg2vec/models/pairwise.py:171-182 and :248-249,275-278), so their
// 1-vs-all sweep cannot share one candidate tile between queries of different relations.  But
// for a FIXED relation r the projected table
//     TransH:  P_r[e] = ent[e] - (ent[e] . w~_r) w~_r,     w~_r = w[r] / max(|w[r]|, 1e-12)
//     TransD:  P_r[e] = ent[e] + (ent[e] . ent_map[e]) rel_map[r]
// turns the model into TransE over tables [P_r, rel]: score_X(h, r, t) == score_TransE(P_r[h], rel[r], P_r[t]).
//     TransR:  P_r[e] = normalize(ent[e]) . M_r   ([N, d_r]; pairwise.py:405-413,430-442), to be used
//              with the once-normalised relation rows rel^[r] = normalize(rel[r]) (normalize_rows_kernel):
//              TransE over [P_r, rel^] then applies the reference's SECOND normalisation (:463-465).
// This kernel writes P_r with EXACTLY the arithmetic score_group<KGE_TRANSH / KGE_TRANSD / KGE_TRANSR>
// applies to the h / t rows of a triple (same lane -> chunk ownership, same fma order, same
// butterfly), so the identity holds bit for bit and a relation-grouped evaluation can run TransE's
// tiled sweep.
#pragma once
#include "kge_models.cuh"

namespace kge {

template <int VEC>
KGE_DEV void st_chunk(float* __restrict__ row, int c, int d, float4 v) {
  const int j = 4 * c;
  if (VEC == 4) {
    *(reinterpret_cast<float4*>(row) + c) = v;
  } else {
    row[j] = v.x;
    if (j + 1 < d) row[j + 1] = v.y;
    if (j + 2 < d) row[j + 2] = v.z;
    if (j + 3 < d) row[j + 3] = v.w;
  }
}

// one 8-lane group per entity row, 32 rows per 256-thread CTA
template <int MODEL, int VEC>
__global__ void __launch_bounds__(256)
project_rows_kernel(ModelParams P, int64_t r, int64_t n, float* __restrict__ out) {
  const int lane = threadIdx.x & 7;
  const int64_t g = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool valid = g < n;
  const int64_t e = valid ? g : n - 1;   // idle groups shadow the last row (shuffles stay group-complete)
  const int d = P.d, nch = (d + 3) >> 2;
  const float* x = P.tab[0] + (size_t)e * d;
  if (MODEL == KGE_TRANSR) {
    // h'_k = sum_j (h_j * ih) M[j,k]: one sequential fma chain per output element, lane l owns the
    // output chunks l, l+8, ... (score_group<KGE_TRANSR>); out rows are d_r wide
    const int dr = P.dr, nchr = (dr + 3) >> 2;
    const float* M = P.tab[2] + (size_t)r * d * dr;
    float* o = out + (size_t)e * dr;
    float sh = 0.f;
    for (int c = lane; c < nch; c += 8) {
      const float4 a = ld_chunk<VEC>(x, c, d);
#pragma unroll
      for (int k = 0; k < 4; ++k) sh = ffma(f4_get(a, k), f4_get(a, k), sh);
    }
    const float ih = inv_norm_from_sumsq(group_sum(sh));
    for (int c = lane; c < nchr; c += 8) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < d; ++j) {
        const float hn = fmul(__ldg(x + j), ih);
        const float4 mrow = ld_chunk<VEC>(M + (size_t)j * dr, c, dr);
#pragma unroll
        for (int k = 0; k < 4; ++k) f4_at(acc, k) = ffma(hn, f4_get(mrow, k), f4_get(acc, k));
      }
      if (valid) st_chunk<VEC>(o, c, dr, acc);
    }
    return;
  }
  float* o = out + (size_t)e * d;
  if (MODEL == KGE_TRANSH) {
    const float* w = P.tab[2] + (size_t)r * d;
    float sw = 0.f;
    for (int c = lane; c < nch; c += 8) {
      const float4 wv = ld_chunk<VEC>(w, c, d);
#pragma unroll
      for (int k = 0; k < 4; ++k) sw = ffma(f4_get(wv, k), f4_get(wv, k), sw);
    }
    const float iw = inv_norm_from_sumsq(group_sum(sw));
    float a = 0.f;
    for (int c = lane; c < nch; c += 8) {
      const float4 wv = ld_chunk<VEC>(w, c, d), xv = ld_chunk<VEC>(x, c, d);
#pragma unroll
      for (int k = 0; k < 4; ++k) a = ffma(f4_get(xv, k), fmul(f4_get(wv, k), iw), a);
    }
    a = group_sum(a);
    for (int c = lane; c < nch; c += 8) {
      const float4 wv = ld_chunk<VEC>(w, c, d), xv = ld_chunk<VEC>(x, c, d);
      float4 ov;
#pragma unroll
      for (int k = 0; k < 4; ++k) f4_at(ov, k) = ffma(-a, fmul(f4_get(wv, k), iw), f4_get(xv, k));
      if (valid) st_chunk<VEC>(o, c, d, ov);
    }
  } else {  // KGE_TRANSD
    const float* xm = P.tab[2] + (size_t)e * d;
    const float* rm = P.tab[3] + (size_t)r * d;
    const float a = group_dot<VEC>(x, xm, d, nch, lane);
    for (int c = lane; c < nch; c += 8) {
      const float4 rv = ld_chunk<VEC>(rm, c, d), xv = ld_chunk<VEC>(x, c, d);
      float4 ov;
#pragma unroll
      for (int k = 0; k < 4; ++k) f4_at(ov, k) = ffma(a, f4_get(rv, k), f4_get(xv, k));
      if (valid) st_chunk<VEC>(o, c, d, ov);
    }
  }
}

// out[i] = row_i * (1 / max(|row_i|, 1e-12)) — F.normalize in the canonical arithmetic (rule 3), one
// 8-lane group per row: TransR's first normalisation of the relation rows (pairwise.py:430-432).
template <int VEC>
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ in, int64_t n, int width, float* __restrict__ out) {
  const int lane = threadIdx.x & 7;
  const int64_t g = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool valid = g < n;
  const int64_t i = valid ? g : n - 1;
  const int nch = (width + 3) >> 2;
  const float* x = in + (size_t)i * width;
  float s = 0.f;
  for (int c = lane; c < nch; c += 8) {
    const float4 b = ld_chunk<VEC>(x, c, width);
#pragma unroll
    for (int k = 0; k < 4; ++k) s = ffma(f4_get(b, k), f4_get(b, k), s);
  }
  const float inv = inv_norm_from_sumsq(group_sum(s));
  for (int c = lane; c < nch; c += 8) {
    const float4 b = ld_chunk<VEC>(x, c, width);
    if (valid) st_chunk<VEC>(out + (size_t)i * width, c, width,
                             make_float4(fmul(b.x, inv), fmul(b.y, inv), fmul(b.z, inv), fmul(b.w, inv)));
  }
}

}  // namespace kge
