// Do DMMA (FP64 tensor) and DFMA (FP64 scalar) share execution resources on B200?  Times DMMA-only, DFMA-only and an
// interleaved mix with the same instruction counts; if mixed time ~= sum of the two, the pipes are shared.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>  // 0 dmma only, 1 dfma only, 2 both
__global__ void k(double* out, int iters, double a, double b) {
  double c[4][2], x[8];
  for (int j = 0; j < 4; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
  for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 0.5 + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (MODE != 1) dmma(c[j][0], c[j][1], a, b);
      if (MODE != 0) { x[2 * j] = fma(x[2 * j], a, b); x[2 * j + 1] = fma(x[2 * j + 1], a, b); }
    }
  }
  double s = 0;
  for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1];
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
float run(double* out, int blocks, int threads, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 2048);
  const int iters = 8000;
  for (int wps : {16, 32}) {
    const int threads = 256, blocks = p.multiProcessorCount * wps * 32 / threads;
    float t0 = run<0>(out, blocks, threads, iters), t1 = run<1>(out, blocks, threads, iters), t2 = run<2>(out, blocks, threads, iters);
    printf("warps/SM %d: dmma-only %.3f ms, dfma-only %.3f ms (4 DMMA : 8 DFMA per iteration), mixed %.3f ms, sum %.3f, max %.3f\n", wps, t0, t1, t2, t0 + t1,
           t0 > t1 ? t0 : t1);
  }
  return 0;
}
