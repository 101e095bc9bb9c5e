// kge_project.cu — C-ABI launcher of the per-relation entity projection (kernel: kge_project.cuh).
#include "kge_project.cuh"

namespace kge {
int check_model(const kge_model_t* m);   // kge_score.cu
}

using namespace kge;

extern "C" int kge_project_entities(const kge_model_t* m, int64_t r, float* out, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  if (m->model != KGE_TRANSH && m->model != KGE_TRANSD) {
    set_error("kge_project_entities: only TransH and TransD project their entity rows per relation");
    return KGE_ENOTSUP;
  }
  if (!out || r < 0 || r >= m->num_rel) { set_error("kge_project_entities: bad arguments"); return KGE_EINVAL; }
  const ModelParams P = make_params(m, nullptr);
  int vec = pick_vec(m, m->model == KGE_TRANSH ? 3 : 4, m->dim);
  if (vec == 4 && ((uintptr_t)out & 15)) vec = 2;
  if (vec == 2 && ((uintptr_t)out & 7)) vec = 1;
  const int64_t n = m->num_ent;
  const unsigned grid = (unsigned)((n + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(M)                                                                       \
  do {                                                                                  \
    if (vec == 4) project_rows_kernel<M, 4><<<grid, 256, 0, st>>>(P, r, n, out);        \
    else if (vec == 2) project_rows_kernel<M, 2><<<grid, 256, 0, st>>>(P, r, n, out);   \
    else project_rows_kernel<M, 1><<<grid, 256, 0, st>>>(P, r, n, out);                 \
  } while (0)
  if (m->model == KGE_TRANSH) LAUNCH(KGE_TRANSH); else LAUNCH(KGE_TRANSD);
#undef LAUNCH
  KGE_CHECK_LAUNCH("project_rows_kernel");
  return KGE_OK;
}
