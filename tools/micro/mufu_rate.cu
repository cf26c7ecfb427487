// MUFU.EX2 issue rate per SM (the softmax stage of the tcgen05 attention kernels is sized against it).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ex2(float* out, int iters, float seed) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed - 0.01f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fma(float* out, int iters, float seed) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed - 0.01f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float* out;
  const int threads = 512, blocks = p.multiProcessorCount * 2, iters = 4096;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      if (which == 0) k_ex2<<<blocks, threads>>>(out, iters, -0.5f);
      else k_fma<<<blocks, threads>>>(out, iters, 0.999f);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      const double ops = (double)threads * blocks * iters * 8;
      printf("%s: %.3f ms, %.1f Gop/s, %.2f ops/clk/SM at the nominal %d MHz\n", which == 0 ? "ex2" : "fma", ms, ops / ms / 1e6,
             ops / (ms * 1e-3) / p.multiProcessorCount / (clk_khz * 1e3), clk_khz / 1000);
    }
  }
  return 0;
}
