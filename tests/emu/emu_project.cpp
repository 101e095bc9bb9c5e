// tests/emu/emu_project.cpp — TEST INFRASTRUCTURE.  Runs project_rows_kernel
// (pykg2vec_b200/csrc/kge_project.cuh) on the host, CUDA thread by CUDA thread
// (tests/emu/cuda_runtime.h); tests/test_emu_project.py checks that TransE over the projected
// table gives TransH's / TransD's scores bit for bit (oracle on both sides).
#include "kge_project.cuh"

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

template <int MODEL>
static void run(const ModelParams& P, int vec, int64_t r, int64_t n, float* out) {
  const dim3 grid((unsigned)((n + 31) / 32)), block(256);
  if (vec == 4) cuda_emu::launch(grid, block, [&] { project_rows_kernel<MODEL, 4>(P, r, n, out); });
  else if (vec == 2) cuda_emu::launch(grid, block, [&] { project_rows_kernel<MODEL, 2>(P, r, n, out); });
  else cuda_emu::launch(grid, block, [&] { project_rows_kernel<MODEL, 1>(P, r, n, out); });
}

extern "C" int emu_project_entities(const kge_model_t* m, int64_t r, float* out) {
  const ModelParams P = make_params(m, nullptr);
  if (m->model == KGE_TRANSR) {
    run<KGE_TRANSR>(P, pick_vec(m, 3, m->dim, m->rel_dim), r, m->num_ent, out);
    return 0;
  }
  const int vec = pick_vec(m, m->model == KGE_TRANSH ? 3 : 4, m->dim);
  if (m->model == KGE_TRANSH) run<KGE_TRANSH>(P, vec, r, m->num_ent, out);
  else if (m->model == KGE_TRANSD) run<KGE_TRANSD>(P, vec, r, m->num_ent, out);
  else return -1;
  return 0;
}

extern "C" int emu_normalize_rows(const float* in, int64_t n, int width, int vec, float* out) {
  const dim3 grid((unsigned)((n + 31) / 32)), block(256);
  if (vec == 4) cuda_emu::launch(grid, block, [&] { normalize_rows_kernel<4>(in, n, width, out); });
  else if (vec == 2) cuda_emu::launch(grid, block, [&] { normalize_rows_kernel<2>(in, n, width, out); });
  else cuda_emu::launch(grid, block, [&] { normalize_rows_kernel<1>(in, n, width, out); });
  return 0;
}
