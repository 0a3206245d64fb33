// FP64 FMA / exp throughput micro-benchmark (SURVEY.md section 8d asks for it: MEASURED_PEAKS.json has no
// FP64 figure).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/dfma_bench profiles/dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters) {
  double a[8];
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-9 + k;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
  double s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dexp(double* out, int iters) {
  double a[4];
  for (int k = 0; k < 4; ++k) a[k] = -1e-3 * (threadIdx.x + k);
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = exp(a[k]) - 1.0000001;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3];
}
__global__ void ffma(float* out, int iters) {
  float a[8];
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-9f + k;
  const float b = 1.0000001f, c = 1e-9f;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], b, c);
  float s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* d; cudaMalloc(&d, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); dfma<<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("DFMA: %.2f TFLOP/s (%.1f FMA/clk/SM at %d MHz)\n", 2.0 * blocks * threads * 8.0 * iters / ms / 1e9,
                    (double)blocks * threads * 8.0 * iters / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
    cudaEventRecord(e0); ffma<<<blocks, threads>>>((float*)d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("FFMA: %.2f TFLOP/s\n", 2.0 * blocks * threads * 8.0 * iters / ms / 1e9);
    cudaEventRecord(e0); dexp<<<blocks, threads>>>(d, iters / 10); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("exp(double): %.3e /s\n", (double)blocks * threads * 4.0 * (iters / 10) / (ms * 1e-3));
  }
  return 0;
}
