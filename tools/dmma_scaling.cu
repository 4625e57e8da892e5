// DMMA (mma.sync m8n8k4 f64) throughput on B200 as a function of resident warps per SM and independent
// accumulator chains per warp.  Explains how many warps the fused kernel needs to keep the FP64 tensor pipe busy.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CHAINS>
__global__ void k(double* out, int iters, double a, double b) {
  double c[CHAINS][2];
  for (int j = 0; j < CHAINS; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) dmma(c[j][0], c[j][1], a, b);
  }
  double s = 0;
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS>
double run(double* out, int sms, int warps_per_sm) {
  const int threads = 32 * (warps_per_sm < 8 ? warps_per_sm : 8), blocks = sms * (warps_per_sm * 32 / threads);
  const int iters = 8000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CHAINS><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<CHAINS><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return 2.0 * 256 * CHAINS * iters * (double)blocks * (threads / 32) / (best * 1e-3) / 1e12;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 2048);
  printf("warps/SM  chains=1  chains=3  chains=6  chains=8   (FP64 DMMA TFLOP/s)\n");
  for (int w : {4, 8, 16, 32, 64})
    printf("%8d  %8.2f  %8.2f  %8.2f  %8.2f\n", w, run<1>(out, p.multiProcessorCount, w), run<3>(out, p.multiProcessorCount, w),
           run<6>(out, p.multiProcessorCount, w), run<8>(out, p.multiProcessorCount, w));
  return 0;
}
