// kge_loss.cu — loss functions of pykg2vec/utils/criterion.py and the get_reg()
// regularisers of the pointwise models, each as value + gradient in one pass.
// Score vectors are tiny (B*(1+neg) floats): one CTA, deterministic tree reduction.
#include "kge_models.cuh"

namespace kge {

constexpr int kLossThreads = 1024;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  }
  return v;  // valid in thread 0
}

// Criterion.pairwise_hinge, criterion.py:26-29
__global__ void __launch_bounds__(kLossThreads)
hinge_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int64_t n, float margin,
             float* __restrict__ loss, float* __restrict__ gpos, float* __restrict__ gneg) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = fmaxf(fsub(fadd(pos[i], margin), neg[i]), 0.f);
    acc += v;
    if (gpos) gpos[i] = v > 0.f ? 1.f : 0.f;
    if (gneg) gneg[i] = v > 0.f ? -1.f : 0.f;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = acc;
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float logsigmoid_f(float x) { return x < 0.f ? x - log1pf(expf(x)) : -log1pf(expf(-x)); }

// Criterion.pointwise_logistic, criterion.py:32-34 (F.softplus threshold 20)
__global__ void __launch_bounds__(kLossThreads)
logistic_kernel(const float* __restrict__ preds, const float* __restrict__ target, int64_t n,
                float* __restrict__ loss, float* __restrict__ gpreds) {
  __shared__ float red[32];
  float acc = 0.f;
  const float inv_n = 1.f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float y = target[i], x = y * preds[i];
    acc += softplus_f(x);
    if (gpreds) gpreds[i] = (x > 20.f ? 1.f : sigmoid_f(x)) * y * inv_n;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = acc * inv_n;
}

// Criterion.pariwise_logistic, criterion.py:14-23.  One warp per positive; the softmax
// weights are detached (no gradient flows through them).
__global__ void __launch_bounds__(256)
selfadv_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int64_t B, int neg_rate,
               float alpha, float* __restrict__ loss, float* __restrict__ gpos,
               float* __restrict__ gneg) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= B) return;
  const float* ng = neg + i * neg_rate;
  const float inv_b = 1.f / (float)B;
  float mx = -INFINITY;
  for (int j = lane; j < neg_rate; j += 32) mx = fmaxf(mx, -ng[j] * alpha);
#pragma unroll
  for (int off = 16; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float den = 0.f;
  for (int j = lane; j < neg_rate; j += 32) den += expf(-ng[j] * alpha - mx);
#pragma unroll
  for (int off = 16; off; off >>= 1) den += __shfl_xor_sync(0xffffffffu, den, off);
  float row = 0.f;
  for (int j = lane; j < neg_rate; j += 32) {
    const float w = expf(-ng[j] * alpha - mx) / den;
    row += w * logsigmoid_f(ng[j]);
    if (gneg) gneg[i * neg_rate + j] = -inv_b * w * sigmoid_f(-ng[j]);
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) row += __shfl_xor_sync(0xffffffffu, row, off);
  if (lane == 0) {
    const float p = pos[i];
    atomicAdd(loss, -inv_b * (row + logsigmoid_f(-p)));
    if (gpos) gpos[i] = inv_b * sigmoid_f(p);
  }
}

// get_reg(): lmbda * mean_i sum_rows sum_j g(x_j)   (pointwise.py:448-458,190-202,224-238,377-388)
struct GradTablesR { float* t[KGE_MAX_TABLES]; };

KGE_DEV float reg_g(float x, int type) {
  return type == 0 ? x * x : (type == 1 ? x * x * x : fabsf(x) * x * x);
}
KGE_DEV float reg_dg(float x, int type) {
  return type == 0 ? 2.f * x : (type == 1 ? 3.f * x * x : 3.f * x * fabsf(x));
}

template <int MODEL, int VEC>
__global__ void __launch_bounds__(256)
reg_kernel(ModelParams P, GradTablesR GT, int reg_type, float scale, const int64_t* __restrict__ h,
           const int64_t* __restrict__ r, const int64_t* __restrict__ t, int64_t n,
           float* __restrict__ out, float grad_scale, int want_grad) {
  const int lane = threadIdx.x & 7;
  const int64_t g = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool valid = g < n;
  const int64_t gi = valid ? g : n - 1;
  const int64_t hi = __ldg(h + gi), ri = __ldg(r + gi), ti = __ldg(t + gi);
  TripleRows R;
  resolve_rows<MODEL>(R, P, P.tab, P.tab, P.tab, hi, ri, ti);
  // gathered rows, their widths and their gradient rows (up to 24: OctonionE)
  constexpr int kMaxRows = (MODEL == KGE_OCTONIONE) ? 24 : (MODEL == KGE_QUATE ? 12 : 9);
  const float* rows[kMaxRows];
  int width[kMaxRows];
  float* grows[kMaxRows];
#pragma unroll
  for (int k = 0; k < kMaxRows; ++k) { rows[k] = nullptr; width[k] = P.d; grows[k] = nullptr; }
  if (MODEL == KGE_DISTMULT || MODEL == KGE_CP || MODEL == KGE_COMPLEX) {
    rows[0] = R.h[0]; rows[1] = R.h[1]; rows[2] = R.r[0]; rows[3] = R.r[1]; rows[4] = R.t[0]; rows[5] = R.t[1];
  }
  const size_t d = (size_t)P.d;
  auto at = [&](int k, size_t off) -> float* { return (want_grad && valid && GT.t[k]) ? GT.t[k] + off : nullptr; };
  if (MODEL == KGE_DISTMULT) { grows[0] = at(0, hi * d); grows[2] = at(1, ri * d); grows[4] = at(0, ti * d); }
  if (MODEL == KGE_CP) { grows[0] = at(0, hi * d); grows[2] = at(1, ri * d); grows[4] = at(2, ti * d); }
  if (MODEL == KGE_COMPLEX) {
    grows[0] = at(0, hi * d); grows[1] = at(1, hi * d); grows[2] = at(2, ri * d);
    grows[3] = at(3, ri * d); grows[4] = at(0, ti * d); grows[5] = at(1, ti * d);
  }
  if (MODEL == KGE_ANALOGY) {  // (re, im) half-width rows + full-width rows (pointwise.py:106-119)
    const size_t d2 = d / 2;
    rows[0] = R.h[1]; rows[1] = R.h[2]; rows[2] = R.r[1]; rows[3] = R.r[2]; rows[4] = R.t[1]; rows[5] = R.t[2];
    rows[6] = R.h[0]; rows[7] = R.r[0]; rows[8] = R.t[0];
    for (int k = 0; k < 6; ++k) width[k] = P.d / 2;
    grows[0] = at(2, hi * d2); grows[1] = at(3, hi * d2); grows[2] = at(4, ri * d2); grows[3] = at(5, ri * d2);
    grows[4] = at(2, ti * d2); grows[5] = at(3, ti * d2);
    grows[6] = at(0, hi * d); grows[7] = at(1, ri * d); grows[8] = at(0, ti * d);
  }
  if (MODEL == KGE_QUATE || MODEL == KGE_OCTONIONE) {  // pointwise.py:696-727 / :901-960
    constexpr int C = (MODEL == KGE_QUATE) ? 4 : 8;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      rows[c] = R.h[c]; rows[C + c] = R.t[c]; rows[2 * C + c] = R.r[c];
      grows[c] = at(c, hi * d); grows[C + c] = at(c, ti * d); grows[2 * C + c] = at(C + c, ri * d);
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxRows; ++k) {
    if (!rows[k]) continue;
    const int w = width[k], nchk = (w + 3) >> 2;
    for (int c = lane; c < nchk; c += 8) {
      const float4 v = ld_chunk<VEC>(rows[k], c, w);
      float4 gv;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc += reg_g(f4_get(v, e), reg_type);
        f4_at(gv, e) = grad_scale * scale * reg_dg(f4_get(v, e), reg_type);
      }
      if (grows[k]) red_chunk<VEC>(grows[k], c, w, gv);
    }
  }
  // block reduction -> one atomic per CTA
  __shared__ float red[8];
  if (!valid) acc = 0.f;
#pragma unroll
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(out, s * scale);
  }
}

// Rescal.get_normalized_data pairwise.py:862-865: row / ||row||_2, in place, no epsilon.
// One warp per row (rows of the relation-matrix table are d*d wide).
__global__ void __launch_bounds__(256)
normalize_rows_kernel(float* __restrict__ table, int64_t rows, int64_t width) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* x = table + row * width;
  // canonical RSUM: element j -> partial (j>>2)&7; lanes 0..7 own the partials, the other 24 lanes
  // help by handling the same partial for later chunks: lane l takes chunks c = l, l+32, ... whose
  // partial index is l&7; per-partial order must stay increasing in j, so partial p is summed by
  // lanes p, p+8, p+16, p+24 over disjoint, INTERLEAVED chunk sets -> not the canonical order.
  // Canonical order therefore uses 8 lanes per row; the remaining lanes only help with the scaling.
  float s = 0.f;
  if (lane < 8) {
    const int64_t nch = (width + 3) >> 2;
    for (int64_t c = lane; c < nch; c += 8) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t j = 4 * c + e;
        if (j < width) s = ffma(x[j], x[j], s);
      }
    }
  }
  s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 4));
  s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 2));
  s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 1));
  s = __shfl_sync(0xffffffffu, s, 0);
  const float inv = __frcp_rn(__fsqrt_rn(s));
  for (int64_t j = lane; j < width; j += 32) x[j] = fmul(x[j], inv);
}

int check_model(const kge_model_t* m);
int model_vec(const kge_model_t* m);

}  // namespace kge

using namespace kge;

extern "C" int kge_normalize_rows(float* table, int64_t rows, int64_t width, void* stream) {
  if (!table || rows <= 0 || width <= 0) { set_error("kge_normalize_rows: bad arguments"); return KGE_EINVAL; }
  normalize_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(table, rows, width);
  KGE_CHECK_LAUNCH("normalize_rows_kernel");
  return KGE_OK;
}

extern "C" int kge_loss_pairwise_hinge(const float* pos, const float* neg, int64_t n, float margin,
                                       float* loss_out, float* grad_pos, float* grad_neg, void* stream) {
  if (n < 0 || !loss_out || (n > 0 && (!pos || !neg))) { set_error("kge_loss_pairwise_hinge: bad arguments"); return KGE_EINVAL; }
  hinge_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(pos, neg, n, margin, loss_out, grad_pos, grad_neg);
  KGE_CHECK_LAUNCH("hinge_kernel");
  return KGE_OK;
}

extern "C" int kge_loss_pointwise_logistic(const float* preds, const float* target, int64_t n,
                                           float* loss_out, float* grad_preds, void* stream) {
  if (n <= 0 || !loss_out || !preds || !target) { set_error("kge_loss_pointwise_logistic: bad arguments"); return KGE_EINVAL; }
  logistic_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream>>>(preds, target, n, loss_out, grad_preds);
  KGE_CHECK_LAUNCH("logistic_kernel");
  return KGE_OK;
}

extern "C" int kge_loss_selfadv(const float* pos, const float* neg, int64_t B, int32_t neg_rate,
                                float alpha, float* loss_out, float* grad_pos, float* grad_neg,
                                void* stream) {
  if (B <= 0 || neg_rate <= 0 || !loss_out || !pos || !neg) { set_error("kge_loss_selfadv: bad arguments"); return KGE_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  KGE_CUDA_OK(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  const unsigned grid = (unsigned)((B + 7) / 8);
  selfadv_kernel<<<grid, 256, 0, st>>>(pos, neg, B, neg_rate, alpha, loss_out, grad_pos, grad_neg);
  KGE_CHECK_LAUNCH("selfadv_kernel");
  return KGE_OK;
}

extern "C" int kge_reg_fwd_bwd(const kge_model_t* m, int reg_type, float lmbda, const int64_t* h,
                               const int64_t* r, const int64_t* t, int64_t n, float* reg_out,
                               float grad_scale, float* const* grad_tables, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  if (n <= 0 || !h || !r || !t || !reg_out || reg_type < 0 || reg_type > 2) { set_error("kge_reg_fwd_bwd: bad arguments"); return KGE_EINVAL; }
  const bool hyper = (m->model == KGE_QUATE || m->model == KGE_OCTONIONE);
  if (m->model != KGE_DISTMULT && m->model != KGE_COMPLEX && m->model != KGE_CP && m->model != KGE_ANALOGY && !hyper) {
    set_error("kge_reg_fwd_bwd: model %d has no row regulariser", m->model); return KGE_ENOTSUP;
  }
  const ModelParams P = make_params(m, nullptr);
  GradTablesR GT;
  int vec = model_vec(m);
  for (int k = 0; k < KGE_MAX_TABLES; ++k) {
    GT.t[k] = grad_tables ? grad_tables[k] : nullptr;
    if (GT.t[k]) {
      const uintptr_t a = (uintptr_t)GT.t[k];
      if (vec == 4 && (a & 15)) vec = 2;
      if (vec == 2 && (a & 7)) vec = 1;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  KGE_CUDA_OK(cudaMemsetAsync(reg_out, 0, sizeof(float), st));
  const unsigned grid = (unsigned)((n + 31) / 32);
  // QuatE / OctonionE average over the width as well (torch.mean over [b, d] per gathered table)
  const float scale = hyper ? lmbda / ((float)n * (float)m->dim) : lmbda / (float)n;
  const int want = grad_tables ? 1 : 0;
#define CALL(M, V) reg_kernel<M, V><<<grid, 256, 0, st>>>(P, GT, reg_type, scale, h, r, t, n, reg_out, grad_scale, want)
  switch (m->model) {
    case KGE_DISTMULT: KGE_DISPATCH_VEC(KGE_DISTMULT, vec, CALL); break;
    case KGE_CP: KGE_DISPATCH_VEC(KGE_CP, vec, CALL); break;
    case KGE_ANALOGY: KGE_DISPATCH_VEC(KGE_ANALOGY, vec, CALL); break;
    case KGE_QUATE: KGE_DISPATCH_VEC(KGE_QUATE, vec, CALL); break;
    case KGE_OCTONIONE: KGE_DISPATCH_VEC(KGE_OCTONIONE, vec, CALL); break;
    default: KGE_DISPATCH_VEC(KGE_COMPLEX, vec, CALL); break;
  }
#undef CALL
  KGE_CHECK_LAUNCH("reg_kernel");
  return KGE_OK;
}
