// Instruction-fetch microbenchmark: 16 warps per SM run the SAME amount of arithmetic either as one long
// straight-line body (BODY instructions, no reuse -> every line comes from L2 once per pass, like one window of
// the window kernel) or as a short loop that fits the instruction caches.  Prints cycles per pass.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/icache_bench profiles/icache_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define R4(x) x x x x
#define R16(x) R4(R4(x))
#define R64(x) R4(R16(x))
#define R256(x) R4(R64(x))
#define R1024(x) R4(R256(x))
#define STEP a = fmaf(a, c, d); b = fmaf(b, c, d); e = fmaf(e, c, d); f = fmaf(f, c, d);
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, float c, float d, int passes) {
  float a = threadIdx.x, b = a + 1, e = a + 2, f = a + 3;
  __syncthreads();
  const long long t0 = clock64();
  for (int p = 0; p < passes; ++p) {
    if (MODE == 0) {            // 8192 FFMA straight line = 128 KB
      R1024(STEP) R1024(STEP)
    } else if (MODE == 1) {     // same work, body of 256 FFMA (4 KB) x 32
#pragma unroll 1
      for (int i = 0; i < 32; ++i) { R64(STEP) }
    } else if (MODE == 2) {     // 2048 FFMA straight line = 32 KB, x 4
#pragma unroll 1
      for (int i = 0; i < 4; ++i) { R256(STEP) R256(STEP) }
    } else {                    // 1024 FFMA = 16 KB, x 8
#pragma unroll 1
      for (int i = 0; i < 8; ++i) { R256(STEP) }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = (t1 - t0) / passes;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + e + f;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 64);
  for (int rep = 0; rep < 2; ++rep) {
    k<0><<<148, 512>>>(out, cyc, 1.0001f, 0.5f, 50);
    k<1><<<148, 512>>>(out, cyc, 1.0001f, 0.5f, 50);
    k<2><<<148, 512>>>(out, cyc, 1.0001f, 0.5f, 50);
    k<3><<<148, 512>>>(out, cyc, 1.0001f, 0.5f, 50);
  }
  long long h[4];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("8192 FFMA per warp per pass, 16 warps/SM (issue-bound floor: 32768 cycles)\n");
  printf("  128 KB straight line      : %lld cycles/pass\n", h[0]);
  printf("  4 KB body x 32            : %lld cycles/pass\n", h[1]);
  printf("  32 KB body x 4            : %lld cycles/pass\n", h[2]);
  printf("  16 KB body x 8            : %lld cycles/pass\n", h[3]);
  return 0;
}
