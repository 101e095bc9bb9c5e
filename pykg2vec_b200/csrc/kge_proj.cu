// kge_proj.cu — C-ABI launchers of the projection-model tail (kernels and launch plans:
// kge_proj.cuh) and of ConvE's inference trunk (kge_conve.cuh, which reuses the same tiled GEMM —
// one translation unit so the kernel template is instantiated once).  Declarations and reference
// citations: include/kge_b200.h.
#include <cstdlib>

#include "kge_conve.cuh"

using namespace kge;

namespace {

int check_dims(const char* fn, int64_t B, int64_t N, int32_t k) {
  if (B < 0 || N < 0 || k <= 0 || B > (1ll << 30) || N > (1ll << 30)) {
    set_error("%s: bad shape B=%lld N=%lld k=%d", fn, (long long)B, (long long)N, (int)k);
    return KGE_EINVAL;
  }
  if ((B + PBM - 1) / PBM > 65535) {
    set_error("%s: %lld rows exceed 65535 row tiles per call", fn, (long long)B);
    return KGE_EINVAL;
  }
  return KGE_OK;
}

template <int EPI>
int launch_gemm(const ProjLaunch& L, cudaStream_t st, const char* what) {
  const dim3 grid(L.gx, L.gy, L.gz);
  switch (L.tile) {
    case PROJ_TILE_128x128: proj_gemm_kernel<EPI, 8, 8><<<grid, PTHREADS, 0, st>>>(L.g); break;
    case PROJ_TILE_64x128: proj_gemm_kernel<EPI, 4, 8><<<grid, PTHREADS, 0, st>>>(L.g); break;
    default: proj_gemm_kernel<EPI, 4, 4><<<grid, PTHREADS, 0, st>>>(L.g); break;
  }
  KGE_CHECK_LAUNCH(what);
  return KGE_OK;
}

// CTA tile of the forward / counting launches: proj_pick_tile(), or KGE_PROJ_TILE=0|1|2 in the
// environment (64x64 | 64x128 | 128x128; used by bench_proj.py and the tests to time / check each).
int pick_tile(int64_t M, int64_t N) {
  if (const char* e = getenv("KGE_PROJ_TILE"))
    if (e[0] >= '0' && e[0] <= '2' && e[1] == 0) return e[0] - '0';
  return proj_pick_tile(M, N, sm_count());
}

}  // namespace

extern "C" {

int kge_proj_tail_fwd(const float* x, const float* ent, const float* bias, int64_t B, int64_t N,
                      int32_t k, float* preds, void* stream) {
  if (!x || !ent || !preds) { set_error("kge_proj_tail_fwd: null pointer"); return KGE_EINVAL; }
  if (int rc = check_dims("kge_proj_tail_fwd", B, N, k)) return rc;
  if (B == 0 || N == 0) return KGE_OK;
  return launch_gemm<EPI_STORE>(proj_plan_fwd(x, ent, bias, B, N, k, preds, pick_tile(B, N)),
                                (cudaStream_t)stream, "proj_gemm_kernel<sigmoid>");
}

int kge_proj_tail_bwd(const float* grad_preds, const float* preds, const float* x, const float* ent,
                      int64_t B, int64_t N, int32_t k, float* grad_x, float* grad_ent,
                      float* grad_bias, void* stream) {
  if (!grad_preds || !preds || !x || !ent) { set_error("kge_proj_tail_bwd: null pointer"); return KGE_EINVAL; }
  if (int rc = check_dims("kge_proj_tail_bwd", B, N, k)) return rc;
  if ((N + PBM - 1) / PBM > 65535) { set_error("kge_proj_tail_bwd: N too large"); return KGE_EINVAL; }
  if (B == 0 || N == 0) return KGE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_x)
    if (int rc = launch_gemm<EPI_ATOMIC>(proj_plan_grad_x(grad_preds, preds, ent, B, N, k, grad_x, 2 * sm_count()),
                                         st, "proj_gemm_kernel<grad_x>")) return rc;
  if (grad_ent)
    if (int rc = launch_gemm<EPI_ATOMIC>(proj_plan_grad_ent(grad_preds, preds, x, B, N, k, grad_ent), st,
                                         "proj_gemm_kernel<grad_ent>")) return rc;
  if (grad_bias) {
    proj_colsum_kernel<<<proj_tiles(N, 256), 256, 0, st>>>(grad_preds, preds, (int)B, N, grad_bias);
    KGE_CHECK_LAUNCH("proj_colsum_kernel");
  }
  return KGE_OK;
}

int kge_proj_bce(const float* preds, const float* labels, int64_t B, int64_t N, float label_scale,
                 float label_shift, float grad_scale, float* loss_out, float* grad_preds, void* stream) {
  if (!preds || !labels || !loss_out) { set_error("kge_proj_bce: null pointer"); return KGE_EINVAL; }
  if (B <= 0 || N <= 0) { set_error("kge_proj_bce: empty batch"); return KGE_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * N;
  KGE_CUDA_OK(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  proj_bce_kernel<<<proj_bce_blocks(n, sm_count()), 256, 0, st>>>(
      preds, labels, n, label_scale, label_shift, proj_bce_grad_factor(grad_scale, B, N),
      proj_bce_inv_count(B, N), loss_out, grad_preds);
  KGE_CHECK_LAUNCH("proj_bce_kernel");
  return KGE_OK;
}

int kge_proj_labels(const int64_t* rows, const int64_t* ptr, const int64_t* idx, int64_t B, int64_t N,
                    float* labels, void* stream) {
  if (!ptr || !idx || !labels) { set_error("kge_proj_labels: null pointer"); return KGE_EINVAL; }
  if (B < 0 || N <= 0 || B > 0x7fffffffll) { set_error("kge_proj_labels: bad shape"); return KGE_EINVAL; }
  if (B == 0) return KGE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  KGE_CUDA_OK(cudaMemsetAsync(labels, 0, sizeof(float) * (size_t)B * (size_t)N, st));
  proj_labels_kernel<<<(unsigned)B, 128, 0, st>>>(rows, ptr, idx, N, labels);
  KGE_CHECK_LAUNCH("proj_labels_kernel");
  return KGE_OK;
}

int64_t kge_proj_rank_workspace_bytes(int64_t Q) { return (Q > 0 ? Q : 1) * (int64_t)sizeof(float); }

int kge_proj_rank(const float* x, const float* ent, const float* bias, int64_t Q, int64_t N, int32_t k,
                  const int64_t* tgt, const int64_t* filt_ptr, const int64_t* filt_idx, int64_t filt_nnz,
                  int32_t direction, int32_t* counts, void* workspace, int64_t workspace_bytes,
                  void* stream) {
  if (!x || !ent || !tgt || !counts || !workspace) { set_error("kge_proj_rank: null pointer"); return KGE_EINVAL; }
  if (direction != 0 && direction != 1) { set_error("kge_proj_rank: direction must be 0 or 1"); return KGE_EINVAL; }
  if (int rc = check_dims("kge_proj_rank", Q, N, k)) return rc;
  if (workspace_bytes < kge_proj_rank_workspace_bytes(Q)) {
    set_error("kge_proj_rank: workspace too small");
    return KGE_EWORKSPACE;
  }
  if (Q == 0 || N == 0) return KGE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* thr = (float*)workspace;
  proj_target_kernel<<<proj_tiles(Q, 128), 128, 0, st>>>(x, ent, bias, tgt, (int)Q, k, thr);
  KGE_CHECK_LAUNCH("proj_target_kernel");
  if (int rc = launch_gemm<EPI_COUNT>(proj_plan_count(x, ent, bias, Q, N, k, thr, counts, direction,
                                                      pick_tile(Q, N)), st, "proj_gemm_kernel<count>")) return rc;
  if (filt_ptr && filt_idx && filt_nnz > 0) {
    proj_filter_kernel<<<(unsigned)Q, 128, 0, st>>>(x, ent, bias, tgt, filt_ptr, filt_idx, k, thr, counts,
                                                    2 * direction);
    KGE_CHECK_LAUNCH("proj_filter_kernel");
  }
  return KGE_OK;
}


// ---- ConvE inference trunk ----------------------------------------------------------------

int64_t kge_conve_trunk_workspace_bytes(const kge_conve_t* p, int64_t Q) {
  if (!p || p->hidden_size_1 < 3 || p->hidden_size < p->hidden_size_1) return 0;
  const int h2 = p->hidden_size / p->hidden_size_1;
  const long long F = conve_feat_width(h2, p->hidden_size_1), q = Q > 0 ? Q : 1;
  return (q * F + (long long)conve_fc_slices(F) * q * p->hidden_size) * (int64_t)sizeof(float);
}

int kge_conve_trunk_fwd(const kge_conve_t* p, const int64_t* e, const int64_t* r, int64_t Q, float* x,
                        void* workspace, int64_t workspace_bytes, void* stream) {
  if (!p || !e || !r || !x || !workspace) { set_error("kge_conve_trunk_fwd: null pointer"); return KGE_EINVAL; }
  const int k = p->hidden_size, h1 = p->hidden_size_1;
  if (k <= 0 || h1 < 3 || k / h1 < 2 || (k / h1) * h1 != k || 2 * k > CONVE_MAX_IMAGE) {
    set_error("kge_conve_trunk_fwd: unsupported image (hidden_size %d, hidden_size_1 %d)", k, h1);
    return KGE_ENOTSUP;
  }
  const float* ptrs[] = {p->ent, p->rel, p->bn0_weight, p->bn0_bias, p->bn0_mean, p->bn0_var, p->conv_weight,
                         p->conv_bias, p->bn1_weight, p->bn1_bias, p->bn1_mean, p->bn1_var, p->fc_weight, p->fc_bias};
  for (const float* q : ptrs)
    if (!q) { set_error("kge_conve_trunk_fwd: null parameter tensor"); return KGE_EINVAL; }
  if (Q < 0 || (Q + PBM - 1) / PBM > 65535) { set_error("kge_conve_trunk_fwd: bad Q"); return KGE_EINVAL; }
  if (workspace_bytes < kge_conve_trunk_workspace_bytes(p, Q)) {
    set_error("kge_conve_trunk_fwd: workspace too small");
    return KGE_EWORKSPACE;
  }
  if (Q == 0) return KGE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int h2 = k / h1;
  ConveFeat f{};
  f.ent = p->ent; f.rel = p->rel; f.e = e; f.r = r; f.k = k; f.h2 = h2; f.h1 = h1;
  f.bn0_w = p->bn0_weight; f.bn0_b = p->bn0_bias; f.bn0_mean = p->bn0_mean; f.bn0_var = p->bn0_var;
  f.bn0_eps = p->bn0_eps;
  f.conv_w = p->conv_weight; f.conv_b = p->conv_bias;
  f.bn1_w = p->bn1_weight; f.bn1_b = p->bn1_bias; f.bn1_mean = p->bn1_mean; f.bn1_var = p->bn1_var;
  f.bn1_eps = p->bn1_eps;
  f.feat = (float*)workspace;
  conve_feature_kernel<<<(unsigned)Q, CONVE_THREADS, 0, st>>>(f);
  KGE_CHECK_LAUNCH("conve_feature_kernel");
  const long long F = conve_feat_width(h2, h1);
  float* partial = f.feat + Q * F;
  if (int rc = launch_gemm<EPI_STORE>(conve_plan_fc(f.feat, p->fc_weight, Q, F, k, partial), st,
                                      "proj_gemm_kernel<fc>")) return rc;
  conve_fc_combine_kernel<<<proj_tiles(Q * k, 256), 256, 0, st>>>(partial, conve_fc_slices(F), Q * k, k,
                                                                 p->fc_bias, x);
  KGE_CHECK_LAUNCH("conve_fc_combine_kernel");
  return KGE_OK;
}

}  // extern "C"
