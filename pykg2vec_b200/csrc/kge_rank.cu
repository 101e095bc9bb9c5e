// kge_rank.cu — 1-vs-all link-prediction rank counts: replaces Evaluator.test
// (pykg2vec/utils/evaluator.py:309-334) together with the Python rank walk of
// MetricCalculator.get_tail_rank/get_head_rank (evaluator.py:70-123).
//
// Instead of materialising N scores, sorting them (topk k=N), copying the ids to
// the host and walking the list, each query's rank is COUNTED on the device:
//   raw      = #{e in [row_lo,row_hi) : score_e < score_target}
//   filtered = raw - #{e in filter(q), e != target : score_e < score_target}
// Scores use exactly the arithmetic of kge_score_fwd, so counts of
// disjoint row shards add up to the global rank.
//
// Two sweep implementations:
//   * gather sweep (all models): the fused gather+score group function evaluated
//     over (query, candidate) pairs; one CTA = one query x 512 consecutive candidates.
//   * tiled sweep (kge_rank_tiled.cu; TransE/TransM, DistMult/CP, ComplEx, RotatE):
//     query vectors and candidate rows staged in shared memory by bulk-async copies,
//     register-tiled pair evaluation, FMA-pipe bound.
#include "kge_models.cuh"
#include "kge_rank.cuh"

namespace kge {

constexpr int kThreads = 256;
constexpr int kGroupsPerCta = kThreads / 8;
constexpr int kSweepIters = 16;
constexpr int kCandsPerCta = kGroupsPerCta * kSweepIters;

template <int MODEL, int VEC, int GROUPING>
__global__ void __launch_bounds__(kThreads)
sweep_gather_kernel(ModelParams P, const int64_t* __restrict__ qh, const int64_t* __restrict__ qr,
                    const int64_t* __restrict__ qt, const float* __restrict__ thr, int64_t nc,
                    int32_t* __restrict__ counts, int col, int scratch_floats) {
  extern __shared__ float4 smem_f4[];
  __shared__ int block_cnt;
  float* scratch = reinterpret_cast<float*>(smem_f4) + (size_t)(threadIdx.x >> 3) * scratch_floats;
  const int lane = threadIdx.x & 7, grp = threadIdx.x >> 3;
  const int64_t q = blockIdx.y;
  const int64_t h = __ldg(qh + q), r = __ldg(qr + q), t = __ldg(qt + q);
  const float th = __ldg(thr + q);
  if (threadIdx.x == 0) block_cnt = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kCandsPerCta;
  int cnt = 0;
  for (int it = 0; it < kSweepIters; ++it) {
    const int64_t e = base + it * kGroupsPerCta + grp;
    const bool valid = e < nc;
    const int64_t ei = valid ? e : nc - 1;
    TripleRows R;
    if (GROUPING == KGE_GROUP_TAIL) resolve_rows<MODEL>(R, P, P.qtab, P.tab, P.qtab, h, r, ei);
    else resolve_rows<MODEL>(R, P, P.tab, P.qtab, P.qtab, ei, r, t);
    const float s = score_group<MODEL, VEC, GROUPING>(R, P, lane, scratch);
    cnt += (valid && s < th) ? 1 : 0;
  }
  if (lane == 0 && cnt) atomicAdd(&block_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0 && block_cnt) {
    atomicAdd(counts + q * 4 + col, block_cnt);
    atomicAdd(counts + q * 4 + col + 1, block_cnt);
  }
}

// One group per filter entry (q, e): subtract it from the filtered count when it
// outranks the target.  Entries equal to the target or outside the row shard are skipped.
template <int MODEL, int VEC, int GROUPING>
__global__ void __launch_bounds__(kThreads)
filter_correct_kernel(ModelParams P, const int64_t* __restrict__ qh, const int64_t* __restrict__ qr,
                      const int64_t* __restrict__ qt, const int64_t* __restrict__ tgt,
                      const float* __restrict__ thr, const int64_t* __restrict__ ptr,
                      const int64_t* __restrict__ idx, int64_t Q, int64_t nnz, int64_t row_lo,
                      int64_t row_hi, int32_t* __restrict__ counts, int col, int scratch_floats) {
  extern __shared__ float4 smem_f4[];
  float* scratch = reinterpret_cast<float*>(smem_f4) + (size_t)(threadIdx.x >> 3) * scratch_floats;
  const int lane = threadIdx.x & 7;
  const int64_t k = (int64_t)blockIdx.x * kGroupsPerCta + (threadIdx.x >> 3);
  // the true entry count lives in device memory (ptr[Q]); the host may pass a capacity
  // (upper bound) as nnz so that the launch shape can stay fixed inside a CUDA graph
  const int64_t nnz_true = min(nnz, __ldg(ptr + Q));
  if (nnz_true <= 0) return;
  const bool valid = k < nnz_true;
  const int64_t kk = valid ? k : nnz_true - 1;
  int64_t lo = 0, hi = Q;  // largest q with ptr[q] <= kk
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid) <= kk) lo = mid; else hi = mid;
  }
  const int64_t q = lo;
  const int64_t e = __ldg(idx + kk);
  const bool skip = (e == __ldg(tgt + q)) || e < row_lo || e >= row_hi;
  const int64_t el = skip ? 0 : e - row_lo;
  TripleRows R;
  if (GROUPING == KGE_GROUP_TAIL)
    resolve_rows<MODEL>(R, P, P.qtab, P.tab, P.qtab, __ldg(qh + q), __ldg(qr + q), el);
  else
    resolve_rows<MODEL>(R, P, P.tab, P.qtab, P.qtab, el, __ldg(qr + q), __ldg(qt + q));
  const float s = score_group<MODEL, VEC, GROUPING>(R, P, lane, scratch);
  if (valid && !skip && lane == 0 && s < __ldg(thr + q)) atomicSub(counts + q * 4 + col + 1, 1);
}

// thresholds: the target's own score, evaluated on the query-side tables
template <int MODEL, int VEC, int GROUPING>
__global__ void __launch_bounds__(kThreads)
threshold_kernel(ModelParams P, const int64_t* __restrict__ qh, const int64_t* __restrict__ qr,
                 const int64_t* __restrict__ qt, int64_t Q, float* __restrict__ thr, int scratch_floats) {
  extern __shared__ float4 smem_f4[];
  float* scratch = reinterpret_cast<float*>(smem_f4) + (size_t)(threadIdx.x >> 3) * scratch_floats;
  const int lane = threadIdx.x & 7;
  const int64_t g = (int64_t)blockIdx.x * kGroupsPerCta + (threadIdx.x >> 3);
  const bool valid = g < Q;
  const int64_t gi = valid ? g : Q - 1;
  TripleRows R;
  resolve_rows<MODEL>(R, P, P.qtab, P.qtab, P.qtab, __ldg(qh + gi), __ldg(qr + gi), __ldg(qt + gi));
  const float s = score_group<MODEL, VEC, GROUPING>(R, P, lane, scratch);
  if (valid && lane == 0) thr[g] = s;
}

// Level 2 of the tensor-core sweep (kge_rank_tc.cu) + the filter pass, one kernel: both are exact
// re-evaluations of listed (query, candidate) pairs with the canonical fp32 group function — the
// arithmetic of kge_score_fwd and of the fp32 sweeps.
//   items [0, total)            : pairs whose tensor-core accumulator fell inside the query's band:
//                                 tc_counts[q] += 1 when the candidate really outranks the target
//   items [total, total + nnz)  : filter entries (as filter_correct_kernel): filtered column -= 1
// The fp32 tiled sweep enqueued behind this kernel then commits the direction (counts += tc_counts) or —
// list overflow: ctrl[0] > cap or ctrl[1] — ranks it itself (sweep_tiled_kernel's entry, kge_rank_tiled.cu);
// on overflow this kernel still applies the filter corrections.
template <int MODEL, int VEC, int GROUPING>
__global__ void __launch_bounds__(kThreads)
band_resolve_kernel(ModelParams P, const int64_t* __restrict__ qh, const int64_t* __restrict__ qr,
                    const int64_t* __restrict__ qt, const float* __restrict__ thr,
                    const unsigned long long* __restrict__ list, unsigned* __restrict__ ctrl, unsigned cap,
                    int64_t Q, int32_t* __restrict__ tc_counts, const RankFilter F, int32_t* __restrict__ counts,
                    int col, int scratch_floats) {
  extern __shared__ float4 smem_f4[];
  float* scratch = reinterpret_cast<float*>(smem_f4) + (size_t)(threadIdx.x >> 3) * scratch_floats;
  const int lane = threadIdx.x & 7;
  const unsigned listed = *reinterpret_cast<volatile unsigned*>(&ctrl[0]);
  const bool overflow = listed > cap || *reinterpret_cast<volatile unsigned*>(&ctrl[1]) != 0u;
  const int64_t total = overflow ? 0 : (int64_t)listed;
  const int64_t nnz_true = (F.ptr && F.idx && F.nnz > 0) ? min(F.nnz, __ldg(F.ptr + Q)) : 0;
  const int64_t items = total + nnz_true;
  for (int64_t k = (int64_t)blockIdx.x * kGroupsPerCta + (threadIdx.x >> 3); k < items;
       k += (int64_t)gridDim.x * kGroupsPerCta) {
    int64_t q, e;
    bool skip = false;
    const bool band = k < total;
    if (band) {
      const unsigned long long pr = list[k];
      if (pr == ~0ull) continue;   // unused slot of a warp's reserved block (kge_rank_tc.cu); group-uniform
      q = (int64_t)(pr >> 32);
      e = (int64_t)(pr & 0xffffffffull);
    } else {
      const int64_t kk = k - total;
      int64_t lo = 0, hi = Q;  // largest q with ptr[q] <= kk
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(F.ptr + mid) <= kk) lo = mid; else hi = mid;
      }
      q = lo;
      const int64_t ge = __ldg(F.idx + kk);
      skip = (ge == __ldg(F.tgt + q)) || ge < F.row_lo || ge >= F.row_hi;
      e = skip ? 0 : ge - F.row_lo;
    }
    TripleRows R;
    if (GROUPING == KGE_GROUP_TAIL)
      resolve_rows<MODEL>(R, P, P.qtab, P.tab, P.qtab, __ldg(qh + q), __ldg(qr + q), e);
    else
      resolve_rows<MODEL>(R, P, P.tab, P.qtab, P.qtab, e, __ldg(qr + q), __ldg(qt + q));
    prefetch_triple_rows(R, P.d, P.dr, lane);
    const float s = score_group<MODEL, VEC, GROUPING>(R, P, lane, scratch);
    if (lane == 0 && !skip && s < __ldg(thr + q)) {
      if (band) atomicAdd(tc_counts + q, 1);
      else atomicSub(counts + q * 4 + col + 1, 1);
    }
  }
}

SweepProfile* sweep_profile(int dir) {
  static thread_local SweepProfile prof[2];
  return &prof[dir & 1];
}

int check_model(const kge_model_t* m);
int model_vec(const kge_model_t* m);

int band_resolve(const kge_model_t* m, const kge_model_t* mq, int dir, const int64_t* qh, const int64_t* qr,
                 const int64_t* qt, const float* thr, int64_t Q, const TcDirBuffers& B, const RankFilter& F,
                 int32_t* counts, int col, cudaStream_t st) {
  const ModelParams P = make_params(m, mq);
  int vec = model_vec(m);
  const int vq = model_vec(mq);
  if (vq < vec) vec = vq;
  const int sf = (int)group_scratch_floats(m);
  const size_t smem = (size_t)sf * kGroupsPerCta * sizeof(float);
  const unsigned grid = (unsigned)(2 * sm_count());
#define SET_SMEM_BR(K)                                                                        \
  if (smem > 40 * 1024)                                                                       \
    KGE_CUDA_OK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#define CALL_BR(M, V)                                                                          \
  do {                                                                                         \
    if (dir == 0) { SET_SMEM_BR((band_resolve_kernel<M, V, KGE_GROUP_TAIL>));                  \
      band_resolve_kernel<M, V, KGE_GROUP_TAIL><<<grid, kThreads, smem, st>>>(                 \
          P, qh, qr, qt, thr, B.list, B.ctrl, B.cap, Q, B.tc_counts, F, counts, col, sf); }    \
    else { SET_SMEM_BR((band_resolve_kernel<M, V, KGE_GROUP_HEAD>));                           \
      band_resolve_kernel<M, V, KGE_GROUP_HEAD><<<grid, kThreads, smem, st>>>(                 \
          P, qh, qr, qt, thr, B.list, B.ctrl, B.cap, Q, B.tc_counts, F, counts, col, sf); }    \
  } while (0)
  switch (m->model) {   // the models tc_supported() admits
    case KGE_TRANSE: KGE_DISPATCH_VEC(KGE_TRANSE, vec, CALL_BR); break;
    case KGE_DISTMULT: KGE_DISPATCH_VEC(KGE_DISTMULT, vec, CALL_BR); break;
    case KGE_CP: KGE_DISPATCH_VEC(KGE_CP, vec, CALL_BR); break;
    case KGE_COMPLEX: KGE_DISPATCH_VEC(KGE_COMPLEX, vec, CALL_BR); break;
    case KGE_RESCAL: KGE_DISPATCH_VEC(KGE_RESCAL, vec, CALL_BR); break;
    case KGE_ROTATE: KGE_DISPATCH_VEC(KGE_ROTATE, vec, CALL_BR); break;
    default: set_error("band_resolve: model %d has no tensor-core sweep", (int)m->model); return KGE_ENOTSUP;
  }
#undef CALL_BR
#undef SET_SMEM_BR
  KGE_CHECK_LAUNCH("band_resolve_kernel");
  return KGE_OK;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Fork/join helper: the head-direction chain of a rank call runs on a side stream so that its
// short preparation / filter kernels overlap the other direction's sweep (and fill the idle
// SMs of its last wave).  Streams and events are created lazily, once per host thread and device;
// event record / wait are capture-safe, so the pattern also works inside a CUDA graph capture.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, mid = nullptr, fork2 = nullptr;
  int device = -1;
};
static int side_stream(SideStream** out) {
  static thread_local SideStream ss[16];
  int dev = 0;
  KGE_CUDA_OK(cudaGetDevice(&dev));
  SideStream& s = ss[dev & 15];
  if (!s.stream) {
    KGE_CUDA_OK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    KGE_CUDA_OK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    KGE_CUDA_OK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    KGE_CUDA_OK(cudaEventCreateWithFlags(&s.mid, cudaEventDisableTiming));
    KGE_CUDA_OK(cudaEventCreateWithFlags(&s.fork2, cudaEventDisableTiming));
    s.device = dev;
  }
  *out = &s;
  return KGE_OK;
}

}  // namespace kge

using namespace kge;

extern "C" int64_t kge_rank_workspace_bytes(const kge_model_t* m, int64_t Q) {
  if (!m || Q < 0) return 0;
  size_t bytes = align_up((size_t)2 * (size_t)Q * sizeof(float), 256);
  bytes += tiled_workspace_bytes(m, Q);
  return (int64_t)bytes;
}

extern "C" int kge_rank_1vsall(const kge_model_t* m, const kge_model_t* mq, int64_t row_lo,
                               int64_t row_hi, const int64_t* qh, const int64_t* qr,
                               const int64_t* qt, const int64_t* tgt_h, const int64_t* tgt_t,
                               int64_t Q, const int64_t* filt_t_ptr, const int64_t* filt_t_idx,
                               int64_t filt_t_nnz, const int64_t* filt_h_ptr,
                               const int64_t* filt_h_idx, int64_t filt_h_nnz, int32_t* counts,
                               void* workspace, int64_t workspace_bytes, int flags, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  if (!mq) mq = m;
  rc = check_model(mq);
  if (rc) return rc;
  if (mq->model != m->model || mq->dim != m->dim || mq->rel_dim != m->rel_dim) {
    set_error("kge_rank_1vsall: m and mq describe different models"); return KGE_EINVAL;
  }
  if (Q == 0) return KGE_OK;
  if (Q < 0 || !qh || !qr || !qt || !counts || !workspace || row_lo < 0 || row_hi <= row_lo ||
      row_hi - row_lo > m->num_ent) {
    set_error("kge_rank_1vsall: bad arguments"); return KGE_EINVAL;
  }
  if (Q > 65535) { set_error("kge_rank_1vsall: Q=%lld > 65535, batch the queries", (long long)Q); return KGE_EINVAL; }
  if (workspace_bytes < kge_rank_workspace_bytes(m, Q)) { set_error("workspace too small"); return KGE_EWORKSPACE; }
  if (!tgt_h) tgt_h = qh;
  if (!tgt_t) tgt_t = qt;
  const int64_t nc = row_hi - row_lo;
  cudaStream_t st = (cudaStream_t)stream;
  const ModelParams P = make_params(m, mq);
  int vec = model_vec(m);
  const int vq = model_vec(mq);
  if (vq < vec) vec = vq;
  const int sf = (int)group_scratch_floats(m);
  const size_t smem = (size_t)sf * kGroupsPerCta * sizeof(float);
  if (smem > 227 * 1024) { set_error("kge_rank_1vsall: embedding width too large for this model's scratch"); return KGE_ENOTSUP; }
  float* thr_t = reinterpret_cast<float*>(workspace);
  float* thr_h = thr_t + Q;
  void* tiled_ws = reinterpret_cast<char*>(workspace) + align_up((size_t)2 * (size_t)Q * sizeof(float), 256);
  const unsigned qgrid = (unsigned)((Q + kGroupsPerCta - 1) / kGroupsPerCta);
  const dim3 sgrid((unsigned)((nc + kCandsPerCta - 1) / kCandsPerCta), (unsigned)Q);
  const bool use_tiled = !(flags & KGE_RANK_FORCE_GATHER) && tiled_supported(m);
  const bool use_tc = use_tiled && !(flags & KGE_RANK_NO_TC) && tc_supported(m, nc);
  if (use_tiled) {
    rc = tiled_prepare_candidates(m, nc, tiled_ws, Q, use_tc, st);
    if (rc) return rc;
  }

  for (int d = 0; d < 2; ++d) {
    SweepProfile* sp = sweep_profile(d);
    sp->armed = (flags & KGE_RANK_PROFILE) != 0;
    sp->valid = false;
    if (sp->armed && !sp->beg) {
      KGE_CUDA_OK(cudaEventCreate(&sp->beg));
      KGE_CUDA_OK(cudaEventCreate(&sp->end));
    }
  }
  const bool both = !(flags & (KGE_RANK_HEAD_ONLY | KGE_RANK_TAIL_ONLY));
  // (CP / SimplE refill one shared candidate scratch per direction: their directions stay serial)
  const bool dirs_independent = m->model != KGE_CP && m->model != KGE_SIMPLE && m->model != KGE_SIMPLE_IGNR;
  const bool two_streams = both && use_tiled && dirs_independent && !(flags & KGE_RANK_SINGLE_STREAM);
  SideStream* side = nullptr;
  cudaStream_t main_st = st;
  if (two_streams) {
    rc = side_stream(&side);
    if (rc) return rc;
    KGE_CUDA_OK(cudaEventRecord(side->fork, main_st));          // after candidate preparation
    KGE_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
  }
  // Both directions on the tensor cores: their query preparations run side by side, ONE launch sweeps both
  // (grid.z = 2: same candidate operands; the launch / pipeline-ramp / drain overhead — a third of a 25 us
  // sweep at the FB15k-237 shape — is paid once and the tile units of both directions balance over the SMs),
  // then the two exact-resolution chains run side by side again.
  if (use_tc && two_streams) {
    const RankFilter Ft = {filt_t_ptr, filt_t_idx, filt_t_nnz, tgt_t, row_lo, row_hi};
    const RankFilter Fh = {filt_h_ptr, filt_h_idx, filt_h_nnz, tgt_h, row_lo, row_hi};
    rc = tiled_sweep(m, mq, 0, qh, qr, qt, thr_t, Q, nc, counts, 0, tiled_ws, true, &Ft, nullptr, nullptr, main_st, kSweepPrep);
    if (rc) return rc;
    rc = tiled_sweep(m, mq, 1, qh, qr, qt, thr_h, Q, nc, counts, 2, tiled_ws, true, &Fh, nullptr, nullptr, side->stream, kSweepPrep);
    if (rc) return rc;
    KGE_CUDA_OK(cudaEventRecord(side->mid, side->stream));
    KGE_CUDA_OK(cudaStreamWaitEvent(main_st, side->mid, 0));
    rc = tiled_sweep(m, mq, 0, qh, qr, qt, thr_t, Q, nc, counts, 0, tiled_ws, true, &Ft, nullptr, nullptr, main_st, kSweepTc, true);
    if (rc) return rc;
    KGE_CUDA_OK(cudaEventRecord(side->fork2, main_st));
    KGE_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork2, 0));
    rc = tiled_sweep(m, mq, 0, qh, qr, qt, thr_t, Q, nc, counts, 0, tiled_ws, true, &Ft, nullptr, nullptr, main_st, kSweepPost);
    if (rc) return rc;
    rc = tiled_sweep(m, mq, 1, qh, qr, qt, thr_h, Q, nc, counts, 2, tiled_ws, true, &Fh, nullptr, nullptr, side->stream, kSweepPost);
    if (rc) return rc;
    KGE_CUDA_OK(cudaEventRecord(side->join, side->stream));
    KGE_CUDA_OK(cudaStreamWaitEvent(main_st, side->join, 0));
    return KGE_OK;
  }
  for (int dir = 0; dir < 2; ++dir) {
    if (dir == 0 && (flags & KGE_RANK_HEAD_ONLY)) continue;
    if (dir == 1 && (flags & KGE_RANK_TAIL_ONLY)) continue;
    st = (two_streams && dir == 1) ? side->stream : main_st;
    float* thr = dir == 0 ? thr_t : thr_h;
    const int col = dir == 0 ? 0 : 2;
    const int64_t* fptr = dir == 0 ? filt_t_ptr : filt_h_ptr;
    const int64_t* fidx = dir == 0 ? filt_t_idx : filt_h_idx;
    const int64_t nnz = dir == 0 ? filt_t_nnz : filt_h_nnz;
    const int64_t* tgt = dir == 0 ? tgt_t : tgt_h;
#define SET_SMEM(K)                                                                          \
  if (smem > 40 * 1024)                                                                      \
    KGE_CUDA_OK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#define CALL_THR(M, V)                                                                         \
  do {                                                                                         \
    if (dir == 0) { SET_SMEM((threshold_kernel<M, V, KGE_GROUP_TAIL>));                        \
      threshold_kernel<M, V, KGE_GROUP_TAIL><<<qgrid, kThreads, smem, st>>>(P, qh, qr, qt, Q, thr, sf); } \
    else { SET_SMEM((threshold_kernel<M, V, KGE_GROUP_HEAD>));                                 \
      threshold_kernel<M, V, KGE_GROUP_HEAD><<<qgrid, kThreads, smem, st>>>(P, qh, qr, qt, Q, thr, sf); } \
  } while (0)
    if (!use_tiled) {  // the tiled path computes the thresholds inside its query-prep kernel
      KGE_DISPATCH_MODEL_VEC(m->model, vec, CALL_THR);
      KGE_CHECK_LAUNCH("threshold_kernel");
    }
#undef CALL_THR

    if (use_tiled) {
      const RankFilter F = {fptr, fidx, nnz, tgt, row_lo, row_hi};
      rc = tiled_sweep(m, mq, dir, qh, qr, qt, thr, Q, nc, counts, col, tiled_ws, use_tc, &F, nullptr, nullptr, st);
      if (rc) return rc;
    } else {
#define CALL_SWEEP(M, V)                                                                       \
  do {                                                                                         \
    if (dir == 0) { SET_SMEM((sweep_gather_kernel<M, V, KGE_GROUP_TAIL>));                     \
      sweep_gather_kernel<M, V, KGE_GROUP_TAIL><<<sgrid, kThreads, smem, st>>>(P, qh, qr, qt, thr, nc, counts, col, sf); } \
    else { SET_SMEM((sweep_gather_kernel<M, V, KGE_GROUP_HEAD>));                              \
      sweep_gather_kernel<M, V, KGE_GROUP_HEAD><<<sgrid, kThreads, smem, st>>>(P, qh, qr, qt, thr, nc, counts, col, sf); } \
  } while (0)
      KGE_DISPATCH_MODEL_VEC(m->model, vec, CALL_SWEEP);
#undef CALL_SWEEP
      KGE_CHECK_LAUNCH("sweep_gather_kernel");
    }

    if (fptr && fidx && nnz > 0 && !use_tc) {   // (the tensor-core path's resolve kernel applies the filters)
      const unsigned fgrid = (unsigned)((nnz + kGroupsPerCta - 1) / kGroupsPerCta);
#define CALL_FILT(M, V)                                                                        \
  do {                                                                                         \
    if (dir == 0) { SET_SMEM((filter_correct_kernel<M, V, KGE_GROUP_TAIL>));                   \
      filter_correct_kernel<M, V, KGE_GROUP_TAIL><<<fgrid, kThreads, smem, st>>>(              \
          P, qh, qr, qt, tgt, thr, fptr, fidx, Q, nnz, row_lo, row_hi, counts, col, sf); }     \
    else { SET_SMEM((filter_correct_kernel<M, V, KGE_GROUP_HEAD>));                            \
      filter_correct_kernel<M, V, KGE_GROUP_HEAD><<<fgrid, kThreads, smem, st>>>(              \
          P, qh, qr, qt, tgt, thr, fptr, fidx, Q, nnz, row_lo, row_hi, counts, col, sf); }     \
  } while (0)
      KGE_DISPATCH_MODEL_VEC(m->model, vec, CALL_FILT);
#undef CALL_FILT
      KGE_CHECK_LAUNCH("filter_correct_kernel");
    }
  }
  if (two_streams) {
    KGE_CUDA_OK(cudaEventRecord(side->join, side->stream));
    KGE_CUDA_OK(cudaStreamWaitEvent(main_st, side->join, 0));
  }
  return KGE_OK;
}

// Test / measurement aid for the tensor-core level of one direction: raw accumulators and the two
// per-query thresholds (see include/kge_b200.h).
extern "C" int kge_rank_tc_probe(const kge_model_t* m, const kge_model_t* mq, int64_t row_lo, int64_t row_hi,
                                 const int64_t* qh, const int64_t* qr, const int64_t* qt, int64_t Q, int direction,
                                 float* dots, float* tau, int32_t* counts, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  if (!mq) mq = m;
  rc = check_model(mq);
  if (rc) return rc;
  if (Q <= 0 || Q > 65535 || !qh || !qr || !qt || !counts || !workspace || row_lo < 0 || row_hi <= row_lo ||
      row_hi - row_lo > m->num_ent || (direction != 0 && direction != 1)) {
    set_error("kge_rank_tc_probe: bad arguments"); return KGE_EINVAL;
  }
  const int64_t nc = row_hi - row_lo;
  if (!tiled_supported(m) || !tc_supported(m, nc)) {
    set_error("kge_rank_tc_probe: no tensor-core sweep for this model / table size"); return KGE_ENOTSUP;
  }
  if (workspace_bytes < kge_rank_workspace_bytes(m, Q)) { set_error("workspace too small"); return KGE_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  float* thr = reinterpret_cast<float*>(workspace) + (direction == 0 ? 0 : Q);
  void* tiled_ws = reinterpret_cast<char*>(workspace) + align_up((size_t)2 * (size_t)Q * sizeof(float), 256);
  rc = tiled_prepare_candidates(m, nc, tiled_ws, Q, true, st);
  if (rc) return rc;
  return tiled_sweep(m, mq, direction, qh, qr, qt, thr, Q, nc, counts, direction == 0 ? 0 : 2, tiled_ws, true, nullptr,
                     dots, tau, st);
}

extern "C" int kge_rank_last_sweep_directions(void) {
  SweepProfile* sp = sweep_profile(0);
  return sp->valid ? sp->ndirs : 0;
}

extern "C" int kge_rank_last_sweep_ms(int direction, float* ms) {
  if (!ms || (direction != 0 && direction != 1)) { set_error("kge_rank_last_sweep_ms: bad arguments"); return KGE_EINVAL; }
  SweepProfile* sp = sweep_profile(direction);
  if (!sp->valid) { set_error("kge_rank_last_sweep_ms: the last kge_rank_1vsall of this thread did not profile direction %d", direction); return KGE_EINVAL; }
  KGE_CUDA_OK(cudaEventSynchronize(sp->end));
  KGE_CUDA_OK(cudaEventElapsedTime(ms, sp->beg, sp->end));
  return KGE_OK;
}

// Measurement aid: clock64 stamps of CTA (0,0) of the following tc_sweep_kernel launches are written to
// buf[3 roles][64] (device memory; NULL switches it off).  Roles: 0 TMA producer (slot 0 start, then one
// per acquired stage), 1 MMA issuer (start, queries resident, then per tile: accumulator free, per
// k-block: stage full), 2 epilogue warp (start, per tile: accumulator ready, tile done; slot 63: kernel entry).
extern "C" int kge_debug_set_tc_trace(long long* buf) {
  tc_set_trace(buf);
  return KGE_OK;
}
