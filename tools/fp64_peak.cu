// FP64 peak microbenchmark for B200 (sm_100a): DFMA pipe and DMMA (mma.sync m8n8k4 f64).
// Writes one JSON line. Used to obtain the FP64 roofline denominator (MEASURED_PEAKS.json has none).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
  for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma(c[j][0], c[j][1], a, b);
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  double best_fma = 0, best_mma = 0; int bf_t = 0, bm_t = 0;
  for (int threads : {128, 256, 512, 1024}) {
    int blocks = sms * (2048 / threads);
    float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (tf > best_fma) { best_fma = tf; bf_t = threads; }
    ms = time_ms([&] { dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double tm = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    if (tm > best_mma) { best_mma = tm; bm_t = threads; }
  }
  // sustained (about 2 s) DFMA
  int threads = bf_t, blocks = sms * (2048 / threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  int n = 0; float total = 0;
  while (total < 2000.f) {
    for (int k = 0; k < 20; ++k) dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    n += 20; cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&total, e0, e1);
  }
  double sustained = 2.0 * 8 * iters * (double)blocks * threads * n / (total * 1e-3) / 1e12;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops\": %.2f, \"dfma_threads\": %d, \"fp64_dmma_tflops\": %.2f, "
         "\"dmma_threads\": %d, \"fp64_dfma_tflops_sustained\": %.2f, \"clock_khz\": %d}\n",
         p.name, sms, best_fma, bf_t, best_mma, bm_t, sustained, p.clockRate);
  return 0;
}
