// kge_conve.cuh — ConvE's trunk in inference mode (ConvE.forward / inner_forward,
// pykg2vec/models/projection.py:104-112 and :86-99 with self.training == False: dropouts are
// identities, bn0 / bn1 use their running statistics, bn2 is skipped):
//     image[2*h2, h1] = [ent[e] ; rel[r]]            two gathered rows stacked along the height
//     bn0 -> conv 3x3 (1 -> 32 channels, stride 1, no padding) -> bn1 -> relu -> flatten   (kernel below)
//     x = relu(feat . fc.weight^T + fc.bias)          (the tiled GEMM of kge_proj.cuh over column slices
//                                                      + the ordered combine kernel below)
// Canonical arithmetic (restated by oracle/kge_oracle.c kgeo_conve_trunk_fwd, bit for bit):
//     BatchNorm(v) = fma(v, a, c) with a = w / sqrt(var + eps), c = b - mean * a   (ATen folds the
//     inference transform the same way); conv = 9 sequential fma from 0 in (di, dj) row-major order,
//     then + conv bias; Linear = one sequential fma chain from 0 per slice of CONVE_FC_SLICE consecutive
//     features (so the contraction spreads over many CTAs whatever the batch size), the slice sums added
//     in ascending slice order, then + bias.
#pragma once
#include "kge_proj.cuh"

namespace kge {

constexpr int CONVE_CH = 32;        // conv2d_1 output channels (projection.py:48)
constexpr int CONVE_THREADS = 256;
constexpr int CONVE_MAX_IMAGE = 4096;  // floats of the stacked image held in shared memory (k <= 2048)
constexpr int CONVE_FC_SLICE = 512;    // features per partial product of the Linear layer (multiple of PBK)

struct ConveFeat {
  const float* ent; const float* rel;
  const int64_t* e; const int64_t* r;
  int k, h2, h1;                   // hidden_size, hidden_size_2, hidden_size_1 (k == h2*h1 elements used)
  const float* bn0_w; const float* bn0_b; const float* bn0_mean; const float* bn0_var; float bn0_eps;
  const float* conv_w; const float* conv_b;        // [32,1,3,3], [32]
  const float* bn1_w; const float* bn1_b; const float* bn1_mean; const float* bn1_var; float bn1_eps;
  float* feat;                     // [Q, 32 * (2*h2-2) * (h1-2)]
};

KGE_DEV void bn_fold(float w, float b, float mean, float var, float eps, float& a, float& c) {
  a = __fdiv_rn(w, __fsqrt_rn(fadd(var, eps)));
  c = fsub(b, fmul(mean, a));
}

// one CTA per query
__global__ void __launch_bounds__(CONVE_THREADS) conve_feature_kernel(const ConveFeat p) {
  __shared__ float img[CONVE_MAX_IMAGE];
  __shared__ float wgt[CONVE_CH * 9];
  __shared__ float cb[CONVE_CH], a1[CONVE_CH], c1[CONVE_CH];
  const int q = blockIdx.x, tid = threadIdx.x;
  const int H = 2 * p.h2, W = p.h1, half = p.h2 * p.h1;
  const int Ho = H - 2, Wo = W - 2, plane = Ho * Wo;
  float a0, c0;
  bn_fold(__ldg(p.bn0_w), __ldg(p.bn0_b), __ldg(p.bn0_mean), __ldg(p.bn0_var), p.bn0_eps, a0, c0);
  const float* er = p.ent + __ldg(p.e + q) * (long long)p.k;
  const float* rr = p.rel + __ldg(p.r + q) * (long long)p.k;
  for (int i = tid; i < 2 * half; i += CONVE_THREADS)
    img[i] = ffma(i < half ? __ldg(er + i) : __ldg(rr + (i - half)), a0, c0);
  for (int i = tid; i < CONVE_CH * 9; i += CONVE_THREADS) wgt[i] = __ldg(p.conv_w + i);
  if (tid < CONVE_CH) {
    cb[tid] = __ldg(p.conv_b + tid);
    bn_fold(__ldg(p.bn1_w + tid), __ldg(p.bn1_b + tid), __ldg(p.bn1_mean + tid), __ldg(p.bn1_var + tid),
            p.bn1_eps, a1[tid], c1[tid]);
  }
  __syncthreads();
  float* out = p.feat + (long long)q * (CONVE_CH * plane);
  for (int o = tid; o < CONVE_CH * plane; o += CONVE_THREADS) {
    const int c = o / plane, rem = o - c * plane;
    const int i = rem / Wo, j = rem - i * Wo;
    float acc = 0.f;
#pragma unroll
    for (int di = 0; di < 3; ++di)
#pragma unroll
      for (int dj = 0; dj < 3; ++dj) acc = ffma(wgt[c * 9 + di * 3 + dj], img[(i + di) * W + j + dj], acc);
    out[o] = fmaxf(ffma(fadd(acc, cb[c]), a1[c], c1[c]), 0.f);
  }
}

inline long long conve_feat_width(int h2, int h1) { return (long long)CONVE_CH * (2 * h2 - 2) * (h1 - 2); }

inline int conve_fc_slices(long long F) { return (int)((F + CONVE_FC_SLICE - 1) / CONVE_FC_SLICE); }

// partial[s][Q,k] = feat[Q, slice s] . fc_w[k, slice s]^T   (no bias, no activation)
inline ProjLaunch conve_plan_fc(const float* feat, const float* fc_w, long long Q, long long F, int k,
                                float* partial) {
  ProjLaunch L{};
  ProjGemm& g = L.g;
  g.A = feat; g.sAm = F; g.sAk = 1; g.Ap = nullptr;
  g.B = fc_w; g.sBn = F; g.sBk = 1;
  g.M = (int)Q; g.N = k; g.K = (int)F; g.klen = CONVE_FC_SLICE;
  g.avec = proj_vec_ok(feat, F, 1); g.bvec = proj_vec_ok(fc_w, F, 1);
  g.C = partial; g.ldc = k; g.zstride = Q * k; g.bias = nullptr; g.act = ACT_NONE;
  L.gx = proj_tiles(k, PBN); L.gy = proj_tiles(Q, PBM); L.gz = (unsigned)conve_fc_slices(F);
  return L;
}

// x[i] = relu(((P_0[i] + P_1[i]) + ... + P_{S-1}[i]) + fc_b[i % k])
__global__ void __launch_bounds__(256)
conve_fc_combine_kernel(const float* __restrict__ partial, int slices, long long n, int k,
                        const float* __restrict__ fc_b, float* __restrict__ x) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float acc = __ldg(partial + i);
  for (int s = 1; s < slices; ++s) acc = fadd(acc, __ldg(partial + (long long)s * n + i));
  x[i] = fmaxf(fadd(acc, __ldg(fc_b + (int)(i % k))), 0.f);
}

}  // namespace kge
