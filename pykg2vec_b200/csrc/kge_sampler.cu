// kge_sampler.cu — negative sampling on the device: replaces the CPU sampler processes
// process_function_pairwise / process_function_pointwise (pykg2vec/data/generator.py:42-158):
// for every positive (h,r,t) and each of neg_rate draws, corrupt the tail with probability
// 1 - p_r or the head with probability p_r (p_r = 0.5 for "uniform" sampling, the relation's
// tph/(tph+hpt) for "bern", generator.py:72), re-drawing the replacement entity while the
// corrupted triple is a known positive (generator.py:76-77, 86-87).
//
// The positives live in an open-addressing hash set of packed 64-bit keys in HBM (built once per
// dataset); one thread per negative.  Randomness is a counter-based generator (splitmix64 of
// (seed, step, sample index, attempt)), so a batch is a pure function of its arguments and the
// CPU oracle reproduces it bit-for-bit.  The reference draws from numpy's global Mersenne
// Twister in worker processes — its stream cannot be matched; what is matched is the sampling
// law and the rejection rule.
#include "kge_common.cuh"

namespace kge {

constexpr uint64_t kEmpty = ~0ull;
constexpr int kMaxAttempts = 64;

__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t pack_key(int64_t h, int64_t r, int64_t t) {
  return ((uint64_t)h << 42) | ((uint64_t)r << 22) | (uint64_t)t;  // h,t < 2^22, r < 2^20
}

__global__ void tripleset_insert_kernel(const int64_t* __restrict__ h, const int64_t* __restrict__ r,
                                        const int64_t* __restrict__ t, int64_t n,
                                        unsigned long long* __restrict__ slots, uint64_t mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = pack_key(h[i], r[i], t[i]);
  uint64_t s = mix64(key) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(slots + s, (unsigned long long)kEmpty, (unsigned long long)key);
    if (prev == kEmpty || prev == key) return;
    s = (s + 1) & mask;
  }
}

__device__ __forceinline__ bool tripleset_contains(const uint64_t* __restrict__ slots, uint64_t mask, uint64_t key) {
  uint64_t s = mix64(key) & mask;
  while (true) {
    const uint64_t v = __ldg(slots + s);
    if (v == key) return true;
    if (v == kEmpty) return false;
    s = (s + 1) & mask;
  }
}

// layout 0 (pairwise):  out_* [B*neg_rate], negatives of positive i at [i*neg_rate, (i+1)*neg_rate)
// layout 1 (pointwise): out_* [B*(1+neg_rate)], each positive followed by its negatives; out_y = +1/-1
__global__ void sample_negatives_kernel(const uint64_t* __restrict__ slots, uint64_t mask,
                                        const int64_t* __restrict__ ph, const int64_t* __restrict__ pr,
                                        const int64_t* __restrict__ pt, int64_t B, int neg_rate,
                                        const float* __restrict__ head_prob, int64_t num_ent, uint64_t base,
                                        int layout, int64_t* __restrict__ oh, int64_t* __restrict__ orr,
                                        int64_t* __restrict__ ot, int64_t* __restrict__ oy) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * neg_rate) return;
  const int64_t i = idx / neg_rate, j = idx % neg_rate;
  const int64_t h = ph[i], r = pr[i], t = pt[i];
  const uint64_t s0 = mix64(base + (uint64_t)idx);
  const float u = (float)(s0 >> 40) * (1.0f / 16777216.0f);
  const float prob = head_prob ? __ldg(head_prob + r) : 0.5f;
  const bool corrupt_tail = u > prob;  // generator.py:73 `if np.random.random() > prob`
  int64_t e = 0;
  for (int a = 0; a < kMaxAttempts; ++a) {
    e = (int64_t)__umul64hi(mix64(s0 + (uint64_t)a + 1ull), (uint64_t)num_ent);
    const uint64_t key = corrupt_tail ? pack_key(h, r, e) : pack_key(e, r, t);
    if (!tripleset_contains(slots, mask, key)) break;
  }
  const int64_t o = layout == 0 ? idx : i * (1 + neg_rate) + 1 + j;
  oh[o] = corrupt_tail ? h : e;
  orr[o] = r;
  ot[o] = corrupt_tail ? e : t;
  if (layout == 1) {
    oy[o] = -1;
    if (j == 0) {
      const int64_t p = i * (1 + neg_rate);
      oh[p] = h; orr[p] = r; ot[p] = t; oy[p] = 1;
    }
  }
}

}  // namespace kge

using namespace kge;

extern "C" int64_t kge_tripleset_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

extern "C" int kge_tripleset_build(const int64_t* h, const int64_t* r, const int64_t* t, int64_t n,
                                   uint64_t* slots, int64_t capacity, int64_t num_ent, int64_t num_rel,
                                   void* stream) {
  if (!h || !r || !t || !slots || n < 0 || capacity < 2 * n || (capacity & (capacity - 1))) {
    set_error("kge_tripleset_build: bad arguments (capacity must be a power of two >= 2n)"); return KGE_EINVAL;
  }
  if (num_ent > (1ll << 22) || num_rel > (1ll << 20)) {
    set_error("kge_tripleset_build: key packing supports < 2^22 entities and < 2^20 relations"); return KGE_ENOTSUP;
  }
  cudaStream_t st = (cudaStream_t)stream;
  KGE_CUDA_OK(cudaMemsetAsync(slots, 0xFF, (size_t)capacity * sizeof(uint64_t), st));
  if (n == 0) return KGE_OK;
  tripleset_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
      h, r, t, n, reinterpret_cast<unsigned long long*>(slots), (uint64_t)capacity - 1);
  KGE_CHECK_LAUNCH("tripleset_insert_kernel");
  return KGE_OK;
}

extern "C" int kge_sample_negatives(const uint64_t* slots, int64_t capacity, const int64_t* pos_h,
                                    const int64_t* pos_r, const int64_t* pos_t, int64_t B,
                                    int32_t neg_rate, const float* corrupt_head_prob, int64_t num_ent,
                                    uint64_t seed, uint64_t step, int32_t layout, int64_t* out_h,
                                    int64_t* out_r, int64_t* out_t, int64_t* out_y, void* stream) {
  if (!slots || !pos_h || !pos_r || !pos_t || !out_h || !out_r || !out_t || B <= 0 || neg_rate <= 0 ||
      num_ent <= 0 || (capacity & (capacity - 1)) || (layout != 0 && layout != 1) || (layout == 1 && !out_y)) {
    set_error("kge_sample_negatives: bad arguments"); return KGE_EINVAL;
  }
  const uint64_t base = mix64(seed ^ mix64(step));
  const int64_t n = B * neg_rate;
  sample_negatives_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      slots, (uint64_t)capacity - 1, pos_h, pos_r, pos_t, B, neg_rate, corrupt_head_prob, num_ent, base, layout,
      out_h, out_r, out_t, out_y);
  KGE_CHECK_LAUNCH("sample_negatives_kernel");
  return KGE_OK;
}
