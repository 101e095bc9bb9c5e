// kge_rank.cuh — interface between the rank driver (kge_rank.cu) and the tiled sweep
// (kge_rank_tiled.cu).
#pragma once
#include "kge_common.cuh"

namespace kge {
bool tiled_supported(const kge_model_t* m);
size_t tiled_workspace_bytes(const kge_model_t* m, int64_t Q);
// Once per rank call: candidate-side scratch (normalised / padded copies) shared by both
// directions.
int tiled_prepare_candidates(const kge_model_t* m, int64_t nc, void* ws, int64_t Q, cudaStream_t st);
// dir 0: tail sweep (TAIL grouping), 1: head sweep (HEAD grouping).  Writes the query vectors
// AND the thresholds thr[q] (the target's own score), then adds
// #{e < nc : score(q, e) < thr[q]} to counts[q*4+col] and counts[q*4+col+1].
int tiled_sweep(const kge_model_t* m, const kge_model_t* mq, int dir, const int64_t* qh,
                const int64_t* qr, const int64_t* qt, float* thr, int64_t Q, int64_t nc,
                int32_t* counts, int col, void* ws, cudaStream_t st);
}  // namespace kge
