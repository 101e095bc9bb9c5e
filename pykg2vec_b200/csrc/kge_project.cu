// kge_project.cu — C-ABI launcher of the per-relation entity projection (kernel: kge_project.cuh).
#include "kge_project.cuh"

namespace kge {
int check_model(const kge_model_t* m);   // kge_score.cu
}

using namespace kge;

extern "C" int kge_project_entities(const kge_model_t* m, int64_t r, float* out, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  if (m->model != KGE_TRANSH && m->model != KGE_TRANSD && m->model != KGE_TRANSR) {
    set_error("kge_project_entities: only TransH, TransD and TransR project their entity rows per relation");
    return KGE_ENOTSUP;
  }
  if (!out || r < 0 || r >= m->num_rel) { set_error("kge_project_entities: bad arguments"); return KGE_EINVAL; }
  const ModelParams P = make_params(m, nullptr);
  int vec = m->model == KGE_TRANSR ? pick_vec(m, 3, m->dim, m->rel_dim)
                                   : pick_vec(m, m->model == KGE_TRANSH ? 3 : 4, m->dim);
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
  if (m->model == KGE_TRANSH) LAUNCH(KGE_TRANSH);
  else if (m->model == KGE_TRANSD) LAUNCH(KGE_TRANSD);
  else LAUNCH(KGE_TRANSR);
#undef LAUNCH
  KGE_CHECK_LAUNCH("project_rows_kernel");
  return KGE_OK;
}

extern "C" int kge_normalize_rows_to(const float* table, int64_t rows, int64_t width, float* out, void* stream) {
  if (!table || !out || rows < 0 || width <= 0 || width > 0x7fffffffll) {
    set_error("kge_normalize_rows_to: bad arguments");
    return KGE_EINVAL;
  }
  if (rows == 0) return KGE_OK;
  int vec = (width % 4 == 0) ? 4 : ((width % 2 == 0) ? 2 : 1);
  const uintptr_t a = (uintptr_t)table | (uintptr_t)out;
  if (vec == 4 && (a & 15)) vec = 2;
  if (vec == 2 && (a & 7)) vec = 1;
  const unsigned grid = (unsigned)((rows + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec == 4) normalize_rows_kernel<4><<<grid, 256, 0, st>>>(table, rows, (int)width, out);
  else if (vec == 2) normalize_rows_kernel<2><<<grid, 256, 0, st>>>(table, rows, (int)width, out);
  else normalize_rows_kernel<1><<<grid, 256, 0, st>>>(table, rows, (int)width, out);
  KGE_CHECK_LAUNCH("normalize_rows_kernel");
  return KGE_OK;
}
