// kge_bwd.cu — backward of the fused gather+score: replaces autograd through the
// reference's ATen op chain (loss.backward(), pykg2vec/utils/trainer.py:298) for
// forward() of the models in kge_model_id.  One 8-lane group per triple; row
// gradients are scattered with red.global.add into dense gradient tables.
#include "kge_grads.cuh"

namespace kge {

constexpr int kThreads = 256;
constexpr int kGroupsPerCta = kThreads / 8;

struct GradTables { float* t[KGE_MAX_TABLES]; };

template <int MODEL, int VEC, int CHSEL>
__global__ void __launch_bounds__(kThreads)
score_bwd_kernel(ModelParams P, GradTables GT, const int64_t* __restrict__ h,
                 const int64_t* __restrict__ r, const int64_t* __restrict__ t, int64_t n,
                 const float* __restrict__ gout, int scratch_floats) {
  extern __shared__ float4 smem_f4[];
  float* scratch = reinterpret_cast<float*>(smem_f4) + (size_t)(threadIdx.x >> 3) * scratch_floats;
  const int lane = threadIdx.x & 7;
  const int64_t g = (int64_t)blockIdx.x * kGroupsPerCta + (threadIdx.x >> 3);
  const bool valid = g < n;
  const int64_t gi = valid ? g : n - 1;
  const int64_t hi = __ldg(h + gi), ri = __ldg(r + gi), ti = __ldg(t + gi);
  TripleRows R;
  resolve_rows<MODEL>(R, P, P.tab, P.tab, P.tab, hi, ri, ti);
  GradRows G;
  resolve_grad_rows<MODEL>(G, P, GT.t, hi, ri, ti);
  if (!valid) {  // idle groups run the math (full-warp shuffles) but scatter nothing
#pragma unroll
    for (int c = 0; c < 8; ++c) G.h[c] = G.t[c] = G.r[c] = nullptr;
  }
  if (!valid && (MODEL == KGE_SLM || MODEL == KGE_NTN || MODEL == KGE_SME || MODEL == KGE_SME_BL || MODEL == KGE_CONVKB)) return;
  grad_group<MODEL, VEC, CHSEL>(R, G, P, lane, __ldg(gout + gi), scratch);
}

int check_model(const kge_model_t* m);
int model_vec(const kge_model_t* m);

}  // namespace kge

using namespace kge;

extern "C" int kge_score_bwd(const kge_model_t* m, const int64_t* h, const int64_t* r,
                             const int64_t* t, int64_t n, const float* grad_scores,
                             float* const* grad_tables, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  if (n == 0) return KGE_OK;
  if (n < 0 || !h || !r || !t || !grad_scores || !grad_tables) { set_error("kge_score_bwd: bad arguments"); return KGE_EINVAL; }
  const ModelParams P = make_params(m, nullptr);
  GradTables GT;
  int vec = model_vec(m);
  const int nt = num_tables(m->model);
  for (int k = 0; k < KGE_MAX_TABLES; ++k) {
    GT.t[k] = (k < nt) ? grad_tables[k] : nullptr;
    if (GT.t[k] && k < nt && !(m->model == KGE_TRANSM && k == 2)) {
      const uintptr_t a = (uintptr_t)GT.t[k];
      if (vec == 4 && (a & 15)) vec = 2;
      if (vec == 2 && (a & 7)) vec = 1;
    }
  }
  if (m->model == KGE_TRANSM) GT.t[2] = nullptr;  // theta is a buffer, not a parameter (pairwise.py:315)
  const int sf = (int)group_scratch_floats_bwd(m);
  const size_t smem = (size_t)sf * kGroupsPerCta * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("%s: embedding width too large for this model's per-group scratch (%zu B of shared memory)", "kge_score_bwd", smem);
    return KGE_ENOTSUP;
  }
  const unsigned grid = (unsigned)((n + kGroupsPerCta - 1) / kGroupsPerCta);
  cudaStream_t st = (cudaStream_t)stream;
  const int chsel = (m->model == KGE_TRANSE || m->model == KGE_TRANSM) ? ch_select(m->dim) : 0;
#define LAUNCH(M, V, C)                                                                            \
  do {                                                                                             \
    if (smem > 40 * 1024)                                                                          \
      KGE_CUDA_OK(cudaFuncSetAttribute(score_bwd_kernel<M, V, C>,                                  \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    score_bwd_kernel<M, V, C><<<grid, kThreads, smem, st>>>(P, GT, h, r, t, n, grad_scores, sf);   \
  } while (0)
#define CALL(M, V)                                                       \
  do {                                                                   \
    if (!(M == KGE_TRANSE || M == KGE_TRANSM)) { LAUNCH(M, V, 0); }      \
    else if (chsel == 2) { LAUNCH(M, V, 2); }                            \
    else if (chsel == 4) { LAUNCH(M, V, 4); }                            \
    else if (chsel == 8) { LAUNCH(M, V, 8); }                            \
    else { LAUNCH(M, V, 0); }                                            \
  } while (0)
  KGE_DISPATCH_MODEL_VEC(m->model, vec, CALL);
#undef CALL
#undef LAUNCH
  KGE_CHECK_LAUNCH("score_bwd_kernel");
  return KGE_OK;
}
