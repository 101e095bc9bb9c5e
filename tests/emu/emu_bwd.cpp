// tests/emu/emu_bwd.cpp — TEST INFRASTRUCTURE.  Runs the per-model gradient functions of
// pykg2vec_b200/csrc/kge_grads.cuh (grad_group<MODEL, VEC>: the device math behind kge_score_bwd)
// on the host, CUDA thread by CUDA thread (tests/emu/cuda_runtime.h), with the thread mapping of
// score_bwd_kernel (kge_bwd.cu).  tests/test_emu_score.py compares the accumulated dense
// gradients with the gradients the REFERENCE's own autograd produced (tests/golden/*.npz).
#include "kge_grads.cuh"

namespace cuda_emu {
thread_local dim3 t_threadIdx, t_blockIdx;
dim3 g_blockDim, g_gridDim;
BlockCtx* g_block = nullptr;
std::mutex g_atomic_mu;
}  // namespace cuda_emu

namespace kge {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return KGE_ECUDA; }
void count_launch(int) {}
int sm_count() { return 148; }
int num_tables(int) { return 0; }
}  // namespace kge

using namespace kge;

constexpr int kMaxScratch = 8192;
struct GradTables { float* t[KGE_MAX_TABLES]; };

template <int MODEL, int VEC>
static void bwd_body(ModelParams P, GradTables GT, const int64_t* h, const int64_t* r, const int64_t* t,
                     int64_t n, const float* gout, int scratch_floats) {
  __shared__ __align__(16) float smem[32 * kMaxScratch];
  float* scratch = smem + (size_t)(threadIdx.x >> 3) * scratch_floats;
  const int lane = threadIdx.x & 7;
  const int64_t g = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool valid = g < n;
  const int64_t gi = valid ? g : n - 1;
  const int64_t hi = h[gi], ri = r[gi], ti = t[gi];
  TripleRows R;
  resolve_rows<MODEL>(R, P, P.tab, P.tab, P.tab, hi, ri, ti);
  GradRows G;
  resolve_grad_rows<MODEL>(G, P, GT.t, hi, ri, ti);
  if (!valid) {
    for (int c = 0; c < 8; ++c) G.h[c] = G.t[c] = G.r[c] = nullptr;
  }
  if (!valid && (MODEL == KGE_SLM || MODEL == KGE_NTN || MODEL == KGE_SME || MODEL == KGE_SME_BL || MODEL == KGE_CONVKB)) return;
  grad_group<MODEL, VEC>(R, G, P, lane, gout[gi], scratch);
}

extern "C" int emu_score_bwd(const kge_model_t* m, int vec, int ntab, const int64_t* h, const int64_t* r,
                             const int64_t* t, int64_t n, const float* gout, float* const* grad_tables) {
  const ModelParams P = make_params(m, nullptr);
  GradTables GT;
  for (int k = 0; k < KGE_MAX_TABLES; ++k) GT.t[k] = (k < ntab) ? grad_tables[k] : nullptr;
  if (m->model == KGE_TRANSM) GT.t[2] = nullptr;
  const int sf = (int)group_scratch_floats_bwd(m);
  if (sf > kMaxScratch) return KGE_ENOTSUP;
  const dim3 grid((unsigned)((n + 31) / 32)), block(256);
#define CALL(M, V) cuda_emu::launch(grid, block, [&] { bwd_body<M, V>(P, GT, h, r, t, n, gout, sf); })
  KGE_DISPATCH_MODEL_VEC(m->model, vec, CALL);
#undef CALL
  return KGE_OK;
}
