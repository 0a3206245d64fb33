// k9_metrics.cu -- per-beat error metrics of a reconstruction.
//
// Replaces sample_rmse_loss / sample_prd_loss (functions_ae/loss_functions.py:8-20) and the quality score
// as computed at analysis_performance_traditional_vs_ae.py:178-197:
//   true = float32(median beat), pred = float32(y_gauss + rowmax(median beat))
//   rmse = sqrt(mean((pred - true)^2)), prd = 100 sqrt(sum((true - pred)^2) / sum(true^2)), qs = (320 / 20) / prd
// One warp per beat (n_t <= 320 -> 10 samples per lane): the float32 differences and squares are the reference's;
// the sums over the beat are accumulated in float64 and rounded to float32 once (torch's float32 summation
// order is vectorised and build-dependent; the two agree to ~1e-7 relative, the test bar is 1e-5).
#include "aew_rt.h"
#include <math.h>

namespace {

__global__ void __launch_bounds__(256)
k9_beat_metrics(const double* __restrict__ recon, const double* __restrict__ median, long long n_beats, int n_t,
                double compression_ratio, float* __restrict__ rmse, float* __restrict__ prd, float* __restrict__ qs) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = warp; b < n_beats; b += n_warps) {
    const double* m = median + b * n_t;
    const double* y = recon + b * n_t;
    double mx = -INFINITY;
    bool any_nan = false;
    for (int i = lane; i < n_t; i += 32) {
      const double v = m[i];
      any_nan |= isnan(v);
      mx = fmax(mx, v);
    }
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(AEW_FULL, mx, o));
    if (__any_sync(AEW_FULL, any_nan)) mx = __longlong_as_double(0x7ff8000000000000LL);   // np.max propagates NaN
    double sd = 0.0, st = 0.0;
    for (int i = lane; i < n_t; i += 32) {
      const float t = (float)m[i];
      const float p = (float)__dadd_rn(y[i], mx);
      const float d = __fsub_rn(p, t);
      sd += (double)__fmul_rn(d, d);
      st += (double)__fmul_rn(t, t);
    }
    for (int o = 16; o; o >>= 1) {
      sd += __shfl_xor_sync(AEW_FULL, sd, o);
      st += __shfl_xor_sync(AEW_FULL, st, o);
    }
    if (lane == 0) {
      const float mean_sq = (float)(sd / (double)n_t);
      const float r = sqrtf(mean_sq);
      const float q = __fmul_rn(100.0f, sqrtf((float)sd / (float)st));
      rmse[b] = r;
      prd[b] = q;
      qs[b] = (float)compression_ratio / q;
    }
  }
}

}  // namespace

extern "C" int aew_beat_metrics_f32(const double* d_recon, const double* d_median, long long n_beats, int n_t,
                                    double compression_ratio, float* d_rmse, float* d_prd, float* d_qs,
                                    void* stream) {
  if (n_beats < 0 || n_t <= 0) return AEW_ERR_ARG;
  if (n_beats == 0) return AEW_OK;
  if (!d_recon || !d_median || !d_rmse || !d_prd || !d_qs) return AEW_ERR_ARG;
  long long grid = (n_beats + 7) / 8;
  if (grid > 148LL * 8) grid = 148LL * 8;
  AEW_LAUNCH(k9_beat_metrics, (unsigned)grid, 256, 0, stream, d_recon, d_median, n_beats, n_t, compression_ratio,
             d_rmse, d_prd, d_qs);
  return (int)cudaGetLastError();
}
