// How long does the SM take to start the CTAs of a grid shaped like wg_flow_kernel (128 threads, ~80 registers,
// 37 KB dynamic shared memory, optional 64-column TMEM allocation)?  Records globaltimer at CTA start / end and the
// SM id; prints the spacing of the first-wave starts on an SM and the gap between a CTA's end and its successor.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cta_launch cta_launch.cu ; run: ./cta_launch
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int TMEM>
__global__ void __launch_bounds__(128, 6) probe(unsigned long long* out, int spin_us, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned tbase;
  const unsigned long long t0 = gtimer();
  if (TMEM && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(
        (unsigned)__cvta_generic_to_shared(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long t_a = gtimer();
  float acc = (float)threadIdx.x + (float)(size_t)smem * 0.f;
  while (gtimer() - t0 < (unsigned long long)spin_us * 1000ull) acc = acc * 1.0001f + 1.f;
  if (acc == 12345.f) sink[0] = acc;
  __syncthreads();
  if (TMEM && threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
  }
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    out[4 * blockIdx.x + 0] = t0; out[4 * blockIdx.x + 1] = gtimer(); out[4 * blockIdx.x + 2] = smid;
    out[4 * blockIdx.x + 3] = t_a - t0;
  }
}

template <int TMEM>
void run(const char* name, int smem, int grid) {
  unsigned long long* d; float* sink;
  cudaMalloc(&d, sizeof(unsigned long long) * 4 * grid); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(probe<TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 3; ++rep) probe<TMEM><<<grid, 128, smem>>>(d, 40, sink);
  cudaDeviceSynchronize();
  std::vector<unsigned long long> h(4 * grid);
  cudaMemcpy(h.data(), d, sizeof(unsigned long long) * 4 * grid, cudaMemcpyDeviceToHost);
  unsigned long long base = ~0ull;
  for (int i = 0; i < grid; ++i) base = std::min(base, h[4 * i]);
  // SM 0: sorted starts
  std::vector<double> st, en;
  for (int i = 0; i < grid; ++i) if (h[4 * i + 2] == 0) { st.push_back((h[4 * i] - base) / 1e3); en.push_back((h[4 * i + 1] - base) / 1e3); }
  std::sort(st.begin(), st.end()); std::sort(en.begin(), en.end());
  printf("%-28s smem %5d: SM0 starts [us]:", name, smem);
  for (size_t k = 0; k < st.size() && k < 8; ++k) printf(" %.2f", st[k]);
  double gap = 0; int ng = 0; int resident = 0;
  for (size_t k = 0; k < st.size(); ++k) if (st[k] < en[0]) ++resident;
  for (size_t k = resident; k < st.size(); ++k) { gap += st[k] - en[k - resident]; ++ng; }
  double ta = 0; for (int i = 0; i < grid; ++i) ta += h[4 * i + 3] / 1e3;
  printf(" | resident %d, end->next start gap %.2f us, alloc+barrier %.2f us\n", resident, ng ? gap / ng : 0.0, ta / grid);
  cudaFree(d); cudaFree(sink);
}

int main() {
  const int grid = 148 * 12;
  run<0>("no tmem", 0, grid);
  run<0>("no tmem", 37888, grid);
  run<1>("tmem 64 cols", 0, grid);
  run<1>("tmem 64 cols", 37888, grid);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
