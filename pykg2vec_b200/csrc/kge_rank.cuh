// kge_rank.cuh — interface between the rank driver (kge_rank.cu), the fp32 tiled sweep
// (kge_rank_tiled.cu) and the tensor-core sweep (kge_rank_tc.cu).
#pragma once
#include "kge_common.cuh"

namespace kge {
bool tiled_supported(const kge_model_t* m);
size_t tiled_workspace_bytes(const kge_model_t* m, int64_t Q);
// Once per rank call: candidate-side scratch (normalised / padded copies, and the bf16 split of the
// tensor-core path when `use_tc`) shared by both directions.
int tiled_prepare_candidates(const kge_model_t* m, int64_t nc, void* ws, int64_t Q, bool use_tc, cudaStream_t st);
// dir 0: tail sweep (TAIL grouping), 1: head sweep (HEAD grouping).  Writes the query vectors
// AND the thresholds thr[q] (the target's own score), then adds
// #{e < nc : score(q, e) < thr[q]} to counts[q*4+col] and counts[q*4+col+1].
// use_tc: level 1 on the tensor cores + exact resolution of the ambiguous pairs (kge_rank_tc.cu);
// the fp32 sweep is still enqueued but returns at once unless the pair list overflowed.
// tc_dbg (tests): optional [Q][nc] raw tensor-core accumulators; tc_tau_out: optional [Q][4] band
// coefficients followed by the nc candidate norm bounds.
struct RankFilter;
// phases (bit mask; the driver splits a direction's chain so that ONE tensor-core launch can sweep both
// directions): kSweepPrep = query vectors, thresholds (+ CP's per-direction candidate operands);
// kSweepTc = the tensor-core launch (tc_both: of both directions — call it for dir 0 only);
// kSweepPost = level 2 + the fp32 sweep (the whole sweep when !use_tc, else the flag-gated fallback).
constexpr int kSweepPrep = 1, kSweepTc = 2, kSweepPost = 4, kSweepAll = 7;
int tiled_sweep(const kge_model_t* m, const kge_model_t* mq, int dir, const int64_t* qh,
                const int64_t* qr, const int64_t* qt, float* thr, int64_t Q, int64_t nc,
                int32_t* counts, int col, void* ws, bool use_tc, const RankFilter* filter, float* tc_dbg,
                float* tc_tau_out, cudaStream_t st, int phases = kSweepAll, bool tc_both = false);

// Measurement hook (KGE_RANK_PROFILE): CUDA events recorded immediately around the launch of a
// direction's main sweep kernel (tensor-core or fp32) on the stream it is launched on.
struct SweepProfile { cudaEvent_t beg = nullptr, end = nullptr; bool armed = false, valid = false; int ndirs = 1; };
SweepProfile* sweep_profile(int dir);

// ---- tensor-core sweep (kge_rank_tc.cu) -------------------------------------------------------------
struct TcDirBuffers {
  int32_t* tc_counts;          // [Q] certain counts of level 1 (+ the resolved pairs of level 2)
  unsigned* ctrl;              // [0] pair-list length, [1] overflow ([2], [3] unused)
  unsigned long long* list;    // (q << 32 | local candidate row)
  unsigned cap;
  const float* tau;            // [Q][4] band coefficients (centre, a, b, e) of tc_query_finish
  const float* cn;             // [nc] candidate norm bounds
};
void tc_set_trace(long long* buf);
bool tc_supported(const kge_model_t* m, int64_t nc);
size_t tc_workspace_bytes(const kge_model_t* m, int64_t Q);
// src[k]: the model's own fp32 candidate tables (row pitch m->dim); scratch: optional fp32 copy
// [KC][nc][dp] for the fp32 fallback sweep (normalised for TransE), written by the same kernel
int tc_prepare_candidates(const kge_model_t* m, const float* const src[2], int64_t nc, void* tcws, int64_t Q,
                          float* scratch, cudaStream_t st);
struct TcQueryArgs;
TcQueryArgs tc_query_args(const kge_model_t* m, int dir, void* tcws, int64_t Q);
// ndirs == 2 (dir == 0): both directions in one launch (same candidate operands, grid.z = 2)
int tc_sweep(const kge_model_t* m, int dir, int ndirs, int64_t Q, int64_t nc, void* tcws, float* dbg, cudaStream_t st);
void tc_dir_buffers(const kge_model_t* m, int dir, int64_t Q, void* tcws, TcDirBuffers* out);
// Level 2 (kge_rank.cu): exact fp32 re-evaluation of the listed pairs into tc_counts.  The fp32 sweep
// enqueued next either commits the direction (counts[q*4+col], counts[q*4+col+1] += tc_counts[q]) or, if the
// list overflowed, ranks the whole direction itself.
// The same kernel also applies the direction's filter corrections (the entries of the CSR filter that
// outrank the target are subtracted from the filtered column) — they are exact re-evaluations of
// listed pairs too —, so the tensor-core path needs no separate filter pass.
struct RankFilter { const int64_t* ptr; const int64_t* idx; int64_t nnz; const int64_t* tgt; int64_t row_lo, row_hi; };
int band_resolve(const kge_model_t* m, const kge_model_t* mq, int dir, const int64_t* qh, const int64_t* qr,
                 const int64_t* qt, const float* thr, int64_t Q, const TcDirBuffers& B, const RankFilter& F,
                 int32_t* counts, int col, cudaStream_t st);
}  // namespace kge
