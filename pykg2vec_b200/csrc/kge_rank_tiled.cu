// kge_rank_tiled.cu — shared-memory tiled 1-vs-all sweep (placeholder until implemented).
#include "kge_rank.cuh"
namespace kge {
bool tiled_supported(const kge_model_t*) { return false; }
size_t tiled_workspace_bytes(const kge_model_t*, int64_t) { return 0; }
int tiled_sweep(const kge_model_t*, const kge_model_t*, int, const int64_t*, const int64_t*,
                const int64_t*, const float*, int64_t, int64_t, int32_t*, int, void*, cudaStream_t) {
  set_error("tiled sweep not built");
  return KGE_ENOTSUP;
}
}  // namespace kge
