// Throughput of scalar FFMA against packed FFMA2 (fma.rn.f32x2) on sm_100a: N independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, float a, float b, int iters) {
  float2 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); }
      else v[i] = __ffma2_rn(v[i], a2, b2);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 4, 512>>>(out, 0.999f, 0.001f, iters); else k<1><<<148 * 4, 512>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 4 * 512 * 16.0 * iters;
      printf("%s: %.3f ms, %.1f GFMA/s (%.1f FMA/clk/SM at 1.9 GHz)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e6, fma / ms / 1e6 / 148 / 1.9);
    }
  }
  return 0;
}
