// Latency microbenchmarks used to reason about the window kernel (K3): dependent-issue latency in cycles of the
// instructions its phases are built from, measured by one warp (others idle) unless stated.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/lat_bench profiles/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
// keeps the timed chain between the two clock reads (all are volatile asm, kept in program order)
#define FENCE asm volatile("" : "+d"(a), "+d"(b), "+r"(j), "+r"(u))
__global__ void k(double* out, long long* cyc, double x, int* chase) {
  __shared__ double sm[1024];
  __shared__ int ch[1024];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 1024; i += blockDim.x) { sm[i] = x + i; ch[i] = chase[i]; }
  __syncthreads();
  long long t0, t1;
  double a = x, b = x * 0.5;
  int j = lane;
  unsigned u = lane;
  // 0: dependent DADD
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) a = __dadd_rn(a, x);
  FENCE; t1 = clock64();
  if (tid == 0) cyc[0] = t1 - t0;
  // 1: dependent DFMA
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) a = __fma_rn(a, x, x);
  FENCE; t1 = clock64();
  if (tid == 0) cyc[1] = t1 - t0;
  // 2: dependent DSETP + select (compare-exchange chain)
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) { const bool s = a > b; const double lo = s ? b : a, hi = s ? a : b; a = lo + 0 * hi; b = hi; }
  FENCE; t1 = clock64();
  if (tid == 0) cyc[2] = t1 - t0;
  // 3: dependent double shuffle + DADD
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) a = __dadd_rn(a, __shfl_xor_sync(0xffffffffu, a, 1));
  FENCE; t1 = clock64();
  if (tid == 0) cyc[3] = t1 - t0;
  // 4: dependent LDS (pointer chase)
  j = lane;
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) j = ch[j];
  FENCE; t1 = clock64();
  if (tid == 0) cyc[4] = t1 - t0;
  // 5: __ddiv_rn chain
  FENCE; t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = __ddiv_rn(a, x);
  FENCE; t1 = clock64();
  if (tid == 0) cyc[5] = t1 - t0;
  // 6: __dsqrt_rn chain
  a = fabs(a) + 2.0;
  FENCE; t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = __dsqrt_rn(a + x);
  FENCE; t1 = clock64();
  if (tid == 0) cyc[6] = t1 - t0;
  // 7: __syncthreads round trip, whole block
  __syncthreads();
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) __syncthreads();
  FENCE; t1 = clock64();
  if (tid == 0) cyc[7] = t1 - t0;
  // 8: REDUX (32-bit min) chain
  u = (unsigned)j + lane;
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) u = __reduce_min_sync(0xffffffffu, u + lane);
  FENCE; t1 = clock64();
  if (tid == 0) cyc[8] = t1 - t0;
  // 9: ballot chain
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) u = __ballot_sync(0xffffffffu, (u >> (lane & 7)) & 1u) + i;
  FENCE; t1 = clock64();
  if (tid == 0) cyc[9] = t1 - t0;
  // 10: LDS.64 + DADD chain (load address independent of the sum): throughput of a smem-fed reduction
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) a = __dadd_rn(a, sm[(tid + 8 * i) & 1023]);
  FENCE; t1 = clock64();
  if (tid == 0) cyc[10] = t1 - t0;
  // 11: integer IADD chain
  FENCE; t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) u = u * 3 + i;
  FENCE; t1 = clock64();
  if (tid == 0) cyc[11] = t1 - t0;
  out[tid] = a + b + j + u;
}
int main() {
  double* out; long long* cyc; int* chase;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&chase, 1024 * 4);
  int h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = (i * 37 + 11) & 1023;
  cudaMemcpy(chase, h, sizeof h, cudaMemcpyHostToDevice);
  const char* names[] = {"DADD dep", "DFMA dep", "cex (DSETP+SEL) dep", "shfl.f64+DADD dep", "LDS chase", "ddiv_rn dep",
                         "dsqrt_rn dep", "__syncthreads", "REDUX.min dep", "ballot dep", "LDS+DADD stream", "IMAD dep"};
  for (int nt : {32, 512}) {
    for (int rep = 0; rep < 2; ++rep) k<<<1, nt>>>(out, cyc, 1.000001, chase);
    long long hc[16];
    cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
    printf("threads per block: %d (thread 0's view)\n", nt);
    for (int i = 0; i < 12; ++i) printf("  %-22s %7.1f cycles\n", names[i], (double)hc[i] / N);
  }
  return 0;
}
