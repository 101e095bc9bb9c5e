// tests/emu/emu_score.cpp — TEST INFRASTRUCTURE.  Runs the per-model score functions of
// pykg2vec_b200/csrc/kge_models.cuh (score_group<MODEL, VEC, GROUPING>: the device math behind
// kge_score_fwd and the gather sweep) on the host, CUDA thread by CUDA thread
// (tests/emu/cuda_runtime.h), with the thread -> (triple, lane) mapping of score_fwd_kernel
// (kge_score.cu): 8-lane group per triple, 32 triples per 256-thread CTA, idle groups shadow the
// last triple.  tests/test_emu_score.py compares with the oracle bit for bit for every model, so
// the model math has a CPU regression net that needs no GPU.
#include "kge_models.cuh"

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

constexpr int kMaxScratch = 4096;   // floats of per-group scratch the emulated CTA provides

template <int MODEL, int VEC>
static void score_body(ModelParams P, int grouping, const int64_t* h, const int64_t* r, const int64_t* t,
                       int64_t n, float* out, int scratch_floats) {
  __shared__ __align__(16) float smem[32 * kMaxScratch];
  float* scratch = smem + (size_t)(threadIdx.x >> 3) * scratch_floats;
  const int lane = threadIdx.x & 7;
  const int64_t g = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool valid = g < n;
  const int64_t gi = valid ? g : n - 1;
  TripleRows R;
  resolve_rows<MODEL>(R, P, P.tab, P.tab, P.tab, h[gi], r[gi], t[gi]);
  float s;
  if (grouping == KGE_GROUP_TAIL) s = score_group<MODEL, VEC, KGE_GROUP_TAIL>(R, P, lane, scratch);
  else s = score_group<MODEL, VEC, KGE_GROUP_HEAD>(R, P, lane, scratch);
  if (valid && lane == 0) out[g] = s;
}

extern "C" int emu_score_fwd(const kge_model_t* m, int grouping, int vec, const int64_t* h, const int64_t* r,
                             const int64_t* t, int64_t n, float* out) {
  const ModelParams P = make_params(m, nullptr);
  const int sf = (int)group_scratch_floats(m);
  if (sf > kMaxScratch) return KGE_ENOTSUP;
  const dim3 grid((unsigned)((n + 31) / 32)), block(256);
#define CALL(M, V) cuda_emu::launch(grid, block, [&] { score_body<M, V>(P, grouping, h, r, t, n, out, sf); })
  KGE_DISPATCH_MODEL_VEC(m->model, vec, CALL);
#undef CALL
  return KGE_OK;
}
