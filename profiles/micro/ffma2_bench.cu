// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b){ unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void up(unsigned long long v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
constexpr int ITERS = 4096, ILP = 16;
__global__ void k_scalar(float* out, float a, float b) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-6f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = __fmaf_rn(acc[i], a, b);
  }
  float s = 0; for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b) {
  unsigned long long acc[ILP / 2];
  const unsigned long long a2 = pk(a, a), b2 = pk(b, b);
#pragma unroll
  for (int i = 0; i < ILP / 2; ++i) acc[i] = pk(threadIdx.x * 1e-6f + i, threadIdx.x * 1e-6f + i + 0.5f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a2), "l"(b2));
  }
  float s = 0; for (int i = 0; i < ILP / 2; ++i) { float x, y; up(acc[i], x, y); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k_scalar<<<148 * 8, 256>>>(out, 1.0001f, 0.5f); else k_packed<<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      // scalar: ILP fmas per iter per thread; packed: ILP/2 instr x 2 fmas = ILP fmas per iter per thread
      double fmas = (double)148 * 8 * 256 * ITERS * ILP;
      if (rep == 2) printf("%s: %.3f ms  %.2f T lane-FMA/s\n", mode == 0 ? "FFMA  (scalar)" : "FFMA2 (packed)", ms, fmas / ms / 1e9);
    }
  }
  return 0;
}
