// Micro-benchmark: per-SMSP issue rate of scalar / packed / immediate-form FP32 instructions on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
  const float a2 = a * 1.0001f + threadIdx.x * 1e-9f, b2 = b * 0.999f + threadIdx.x * 1e-9f;  // per-thread registers
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) x[i] = fmaf(x[i], a2, b2);                  // FFMA R, R, R
      if (MODE == 1) x[i] = fmaf(x[i], 0.99993f, 0.00017f);      // FFMA imm
      if (MODE == 2) x[i] = x[i] + b2;                            // FADD
      if (MODE == 3) x[i] = x[i] * a2;                            // FMUL
      if (MODE == 4) x[i] = fmaf(x[i], a, b);                     // FFMA with uniform operands
      if (MODE == 5) x[i] = fminf(x[i], b2) ;                     // FMNMX
    }
    if (MODE == 6) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a2, a2), make_float2(b2, b2));
        x[i] = v.x; x[i + 1] = v.y;
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* d, int n_inst) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, warps = 8;
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 32 * warps>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  double inst = 148.0 * 4 * warps * (double)iters * n_inst;   // warp instructions
  printf("%-22s %.3f ms  %.3f warp-inst/clk/SMSP (at 1.965 GHz)\n", name, ms, inst / (ms * 1e-3) / (148 * 4) / 1.965e9);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  run<0>("FFMA reg,reg,reg", d, 16);
  run<1>("FFMA reg,imm,imm", d, 16);
  run<2>("FADD reg,reg", d, 16);
  run<3>("FMUL reg,reg", d, 16);
  run<4>("FFMA reg,uniform", d, 16);
  run<5>("FMNMX reg,reg", d, 16);
  run<6>("FFMA2 (8 per iter)", d, 8);
  return 0;
}
