// Micro-benchmark: issue / pipe throughput of packed FFMA2 vs scalar FFMA on sm_100a (8 independent chains per thread).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 x[8];
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { x[i].x = fmaf(x[i].x, aa.x, bb.x); x[i].y = fmaf(x[i].y, aa.y, bb.y); }
      else x[i] = __ffma2_rn(x[i], aa, bb);
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 1; warps <= 8; warps *= 2)
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 4, 32 * warps>>>(d, iters, 0.999f, 0.001f); else k<1><<<148 * 4, 32 * warps>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 4 * 32 * warps * (double)iters * 16;
      printf("warps/CTA %d (4 CTAs/SM) %s: %.3f ms, %.1f TFMA/s, %.2f fma/clk/SM\n", warps, mode ? "FFMA2" : "FFMA ", ms, fma / ms * 1e-9,
             fma / (ms * 1e-3) / 148 / 1.965e9);
    }
  return 0;
}
