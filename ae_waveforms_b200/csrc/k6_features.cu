// k6_features.cu -- K6: the 35 np_features of a window from per-beat fiducial indices.
//
// Replaces, per kept 20 s window (data_extract_vitaldb.py:146-162):
//   ecg_features_relative_r          feature_extraction/features_ecg.py:56-86
//   compute_ecg_morphology_features  feature_extraction/features_ecg.py:89-136
//   compute_median_features_ecg      median_sample.py:57-82
// and the flattening of data_create_arrays.py:34-36:
//   row = [x, y of the 10 fiducials | 10 durations | 5 prominences].
//
// One 128-thread block per window.  Phase 1 fills the per-beat relative coordinates
// x = (idx - r) / fs * 1000 (division, then multiplication, as Python evaluates it) and
// y = sig[idx] - sig[r] into shared memory (slot 10 = the inserted 'ECG_R_Peaks' zeros,
// features_ecg.py:111).  Phase 2: each warp takes output columns round-robin, forms the
// per-beat value of its column, drops NaN and selects the median by rank counting
// (<= 72 beats: 72 x 72 comparisons spread over 32 lanes); an even count gives
// (lo + hi) / 2 like np.median.  All arithmetic is single IEEE operations (-fmad=false).
#include "aew_rt.h"
#include <math.h>

#define K6_THREADS 128
#define K6_MAXP 72
#define K6_NFID 10
#define K6_NOUT 35

namespace {

// durations: x[second] - x[first]; slot 10 is the R peak (features_ecg.py:114-125)
__device__ const signed char k6_dur[10][2] = {
    {1, 2},    // P_Duration   = P_Offsets - P_Onsets
    {4, 10},   // Q_Duration   = R_Peaks   - R_Onsets
    {3, 6},    // R_Duration   = S_Peaks   - Q_Peaks
    {10, 5},   // S_Duration   = R_Offsets - R_Peaks
    {8, 9},    // T_Duration   = T_Offsets - T_Onsets
    {1, 10},   // PR_Interval  = R_Peaks   - P_Onsets
    {4, 5},    // QRS_Complex  = R_Offsets - R_Onsets
    {5, 9},    // ST_Interval  = T_Offsets - R_Offsets
    {5, 8},    // ST_Segment   = T_Onsets  - R_Offsets
    {4, 9}};   // QT_Interval  = T_Offsets - R_Onsets
// prominences: y[peak] - nanmax(y[b0], y[b1]) (b1 < 0: single baseline) (features_ecg.py:128-134)
__device__ const signed char k6_prom[5][3] = {
    {0, 1, 2}, {3, 4, -1}, {10, 3, 6}, {6, 5, -1}, {7, 8, 9}};

__device__ __forceinline__ double k6_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// value of output column `col` for beat b; X / Y hold the relative coordinates of fiducial f at [f * stride + b]
__device__ __forceinline__ double k6_value(int col, int b, const double* X, const double* Y, int stride) {
  if (col < 20) return (col & 1) ? Y[(col >> 1) * stride + b] : X[(col >> 1) * stride + b];
  if (col < 30) {
    const int a = k6_dur[col - 20][0], c = k6_dur[col - 20][1];
    return __dsub_rn(X[c * stride + b], X[a * stride + b]);
  }
  const int pk = k6_prom[col - 30][0], b0 = k6_prom[col - 30][1], b1 = k6_prom[col - 30][2];
  double base = Y[b0 * stride + b];
  if (b1 >= 0) {                       // np.nanmax over the two baselines: NaN only if both are NaN
    const double o = Y[b1 * stride + b];
    if (isnan(base)) base = o;
    else if (!isnan(o) && o > base) base = o;
  }
  return __dsub_rn(Y[pk * stride + b], base);
}

// relative coordinates of one fiducial index (features_ecg.py:75-83)
__device__ __forceinline__ void k6_relative_xy(const double* __restrict__ s, int L, int idx, int r, double fs,
                                               double& x, double& y, int* status) {
  if (idx == AEW_FID_NAN) {
    x = y = k6_nan();
  } else if (idx < 0 || idx >= L || r < 0 || r >= L) {
    x = y = k6_nan();
    atomicOr(status, AEW_DEV_FID_RANGE);
  } else {
    x = __dmul_rn(__ddiv_rn((double)(idx - r), fs), 1000.0);
    y = __dsub_rn(s[idx], s[r]);
  }
}

// median of the non-NaN entries of v[0..n) by ONE warp, rank counting (np.nanmedian: even count -> (lo + hi) / 2;
// all NaN -> NaN).  `vals` is scratch for n doubles; the result is returned on every lane.
__device__ __forceinline__ double k6_warp_nanmedian(const double* v, int n, double* vals, double* pick, int lane) {
  int m = 0;
  for (int b0 = 0; b0 < n; b0 += 32) {
    const int b = b0 + lane;
    const double x = b < n ? v[b] : k6_nan();
    const bool ok = b < n && !isnan(x);
    const unsigned mask = __ballot_sync(AEW_FULL, ok);
    if (ok) vals[m + __popc(mask & ((1u << lane) - 1u))] = x;
    m += __popc(mask);
  }
  __syncwarp();
  if (m == 0) return k6_nan();
  const int klo = (m - 1) >> 1, khi = m >> 1;
  for (int i = lane; i < m; i += 32) {
    const double x = vals[i];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const double u = vals[j];
      rank += (u < x) || (u == x && j < i);
    }
    if (rank == klo) pick[0] = x;
    if (rank == khi) pick[1] = x;
  }
  __syncwarp();
  const double lo = pick[0], hi = pick[1];
  __syncwarp();
  return (klo == khi) ? lo : __ddiv_rn(__dadd_rn(lo, hi), 2.0);
}

__global__ void __launch_bounds__(K6_THREADS)
k6_features(const double* __restrict__ sig, const long long* __restrict__ sig_off, long long n_win, int L,
            const int* __restrict__ peaks, const int* __restrict__ n_peaks, int max_peaks,
            const int* __restrict__ fid, const unsigned char* __restrict__ valid, double fs,
            double* __restrict__ out, int* __restrict__ status) {
  __shared__ double X[(K6_NFID + 1) * K6_MAXP];
  __shared__ double Y[(K6_NFID + 1) * K6_MAXP];
  __shared__ double vals[K6_THREADS / 32][K6_MAXP];
  __shared__ double colv[K6_THREADS / 32][K6_MAXP];
  __shared__ double pick[K6_THREADS / 32][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (long long w = blockIdx.x; w < n_win; w += gridDim.x) {
    int nb = n_peaks[w];
    if (nb > max_peaks) nb = max_peaks;
    if (nb > K6_MAXP) nb = K6_MAXP;
    double* row = out + w * K6_NOUT;
    if ((valid && !valid[w]) || nb <= 0) {           // delineation failed: NaN row (data_extract_vitaldb.py:132-135)
      if (tid < K6_NOUT) row[tid] = k6_nan();
      continue;
    }
    const double* s = sig + sig_off[w];
    const int* pk = peaks + w * (long long)max_peaks;
    const int* fw = fid + w * (long long)K6_NFID * max_peaks;
    for (int i = tid; i < (K6_NFID + 1) * nb; i += K6_THREADS) {
      const int f = i / nb, b = i - f * nb;
      double x = 0.0, y = 0.0;
      if (f < K6_NFID) k6_relative_xy(s, L, fw[f * max_peaks + b], pk[b], fs, x, y, status);
      X[f * K6_MAXP + b] = x;
      Y[f * K6_MAXP + b] = y;
    }
    __syncthreads();
    for (int col = warp; col < K6_NOUT; col += K6_THREADS / 32) {
      // per-beat values of this column -> median of the non-NaN ones
      for (int b = lane; b < nb; b += 32) colv[warp][b] = k6_value(col, b, X, Y, K6_MAXP);
      __syncwarp();
      const double m = k6_warp_nanmedian(colv[warp], nb, vals[warp], pick[warp], lane);
      if (lane == 0) row[col] = m;
      __syncwarp();
    }
    __syncthreads();
  }
}

// ---- the three stages as kernels of their own: the per-call drop-ins of features_ecg.py / median_sample.py
// (any number of beats; same arithmetic as the fused kernel above)
__global__ void k6_stage_relative(const double* __restrict__ sig, int L, const int* __restrict__ peaks,
                                  const int* __restrict__ fid, int n_fid, int nb, double fs,
                                  double* __restrict__ xy, int* __restrict__ status) {
  // xy [n_fid][2][nb]: x then y of every fiducial (ecg_features_relative_r, features_ecg.py:56-86)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_fid * nb; i += gridDim.x * blockDim.x) {
    const int f = i / nb, b = i - f * nb;
    double x, y;
    k6_relative_xy(sig, L, fid[f * nb + b], peaks[b], fs, x, y, status);
    xy[(2 * f) * nb + b] = x;
    xy[(2 * f + 1) * nb + b] = y;
  }
}

__global__ void k6_stage_morphology(const double* __restrict__ X, const double* __restrict__ Y, int nb,
                                    double* __restrict__ out) {
  // X, Y [11][nb] (slot 10 = the inserted R peak, zeros); out [15][nb] durations then prominences
  // (compute_ecg_morphology_features, features_ecg.py:89-136)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 15 * nb; i += gridDim.x * blockDim.x) {
    const int c = i / nb, b = i - c * nb;
    out[i] = k6_value(20 + c, b, X, Y, nb);
  }
}

__global__ void __launch_bounds__(K6_THREADS) k6_stage_nanmedian(const double* __restrict__ vals, int rows, int n,
                                                                 double* __restrict__ scratch,
                                                                 double* __restrict__ out) {
  // out[r] = np.nanmedian(vals[r][0..n)) (compute_median_features_ecg, median_sample.py:57-82); a warp per row
  __shared__ double pick[K6_THREADS / 32][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = blockIdx.x * (K6_THREADS / 32) + warp; r < rows; r += gridDim.x * (K6_THREADS / 32)) {
    const double m = k6_warp_nanmedian(vals + (long long)r * n, n, scratch + (long long)r * n, pick[warp], lane);
    if (lane == 0) out[r] = m;
  }
}

// NumPy's pairwise summation by ONE thread (oracle/np_pairwise.py); only for the handful of values of rmssd
__device__ double k6_pairwise(const double* a, long long n) {
  if (n < 8) {
    double r = 0.0;
    for (long long i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
    return r;
  }
  if (n <= 128) {
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    long long i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], a[i + k]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  long long h = n / 2;
  h -= h % 8;
  return __dadd_rn(k6_pairwise(a, h), k6_pairwise(a + h, n - h));
}

// estimate_hr / rmssd of peak lists (remove_samples.py:85-121).  One warp.
//   mode 0 'min': median(diff(min_peaks)); 1 'mean': median of both diff lists; 2 'max': median(diff(max_peaks))
//   out[0] = hr_num / median (hr_num = 60 * fs formed on the host as Python does), NaN for an empty list
//   out[1] = sqrt(mean(diff(diff(max_peaks) / rr_div)^2)), NaN below three peaks
__global__ void __launch_bounds__(32) k6_peak_stats(const long long* __restrict__ min_peaks, int n_min,
                                                    const long long* __restrict__ max_peaks, int n_max, int mode,
                                                    double hr_num, double rr_div, double* __restrict__ scratch,
                                                    double* __restrict__ out) {
  const int lane = threadIdx.x;
  const int nlo = (mode == 2 || n_min < 2) ? 0 : n_min - 1, nhi = (mode == 0 || n_max < 2) ? 0 : n_max - 1;
  const int n = nlo + nhi;
  double* d = scratch;                                   // [n] the interval list
  for (int i = lane; i < n; i += 32)
    d[i] = i < nlo ? (double)(min_peaks[i + 1] - min_peaks[i]) : (double)(max_peaks[i - nlo + 1] - max_peaks[i - nlo]);
  __syncwarp();
  // np.median of integers: the middle value, or the mean of the two middle values
  double lo = 0.0, hi = 0.0;
  const int klo = (n - 1) >> 1, khi = n >> 1;
  for (int i = lane; i < n; i += 32) {
    const double x = d[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (d[j] < x) || (d[j] == x && j < i);
    if (rank == klo) scratch[n] = x;
    if (rank == khi) scratch[n + 1] = x;
  }
  __syncwarp();
  if (lane == 0) {
    lo = scratch[n]; hi = scratch[n + 1];
    const double med = n == 0 ? k6_nan() : (klo == khi ? lo : __ddiv_rn(__dadd_rn(lo, hi), 2.0));
    out[0] = __ddiv_rn(hr_num, med);
    double rms = k6_nan();
    if (n_max >= 3) {
      double* q = scratch + n + 2;                       // [n_max - 2]
      for (int t = 0; t + 2 < n_max; ++t) {
        const double a = __ddiv_rn((double)(max_peaks[t + 1] - max_peaks[t]), rr_div);
        const double b = __ddiv_rn((double)(max_peaks[t + 2] - max_peaks[t + 1]), rr_div);
        const double dd = __dsub_rn(b, a);
        q[t] = __dmul_rn(dd, dd);
      }
      rms = __dsqrt_rn(__ddiv_rn(k6_pairwise(q, n_max - 2), (double)(n_max - 2)));
    }
    out[1] = rms;
  }
}

}  // namespace

extern "C" int aew_fiducial_relative_f64(const double* d_sig, int L, const int* d_peaks, const int* d_fid, int n_fid,
                                         int n_beats, double fs, double* d_xy, int* d_status, void* stream) {
  if (L <= 0 || n_fid < 0 || n_beats < 0 || !(fs > 0)) return AEW_ERR_ARG;
  if (n_fid == 0 || n_beats == 0) return AEW_OK;
  if (!d_sig || !d_peaks || !d_fid || !d_xy || !d_status) return AEW_ERR_ARG;
  AEW_LAUNCH(k6_stage_relative, (unsigned)((n_fid * n_beats + 127) / 128), 128, 0, stream, d_sig, L, d_peaks, d_fid,
             n_fid, n_beats, fs, d_xy, d_status);
  return (int)cudaGetLastError();
}

extern "C" int aew_morphology_f64(const double* d_x, const double* d_y, int n_beats, double* d_out, void* stream) {
  if (n_beats < 0) return AEW_ERR_ARG;
  if (n_beats == 0) return AEW_OK;
  if (!d_x || !d_y || !d_out) return AEW_ERR_ARG;
  AEW_LAUNCH(k6_stage_morphology, (unsigned)((15 * n_beats + 127) / 128), 128, 0, stream, d_x, d_y, n_beats, d_out);
  return (int)cudaGetLastError();
}

extern "C" int aew_nanmedian_rows_f64(const double* d_vals, int rows, int n, double* d_scratch, double* d_out,
                                      void* stream) {
  if (rows < 0 || n < 0) return AEW_ERR_ARG;
  if (rows == 0) return AEW_OK;
  if (!d_out || (n > 0 && (!d_vals || !d_scratch))) return AEW_ERR_ARG;
  AEW_LAUNCH(k6_stage_nanmedian, (unsigned)((rows + 3) / 4), K6_THREADS, 0, stream, d_vals, rows, n, d_scratch, d_out);
  return (int)cudaGetLastError();
}

extern "C" int aew_peak_stats_f64(const long long* d_min_peaks, int n_min, const long long* d_max_peaks, int n_max,
                                  int mode, double hr_num, double rr_div, double* d_scratch, double* d_out,
                                  void* stream) {
  if (n_min < 0 || n_max < 0 || mode < 0 || mode > 2 || !d_scratch || !d_out) return AEW_ERR_ARG;
  if ((n_min > 0 && !d_min_peaks) || (n_max > 0 && !d_max_peaks)) return AEW_ERR_ARG;
  AEW_LAUNCH(k6_peak_stats, 1u, 32, 0, stream, d_min_peaks, n_min, d_max_peaks, n_max, mode, hr_num, rr_div, d_scratch,
             d_out);
  return (int)cudaGetLastError();
}

extern "C" int aew_fiducial_features_f64(const double* d_sig, const long long* d_sig_off, long long n_win, int L,
                                         const int* d_peaks, const int* d_n_peaks, int max_peaks, const int* d_fid,
                                         const unsigned char* d_valid, double fs, double* d_features, int* d_status,
                                         void* stream) {
  if (n_win < 0 || L <= 0 || max_peaks <= 0 || max_peaks > K6_MAXP || !(fs > 0)) return AEW_ERR_ARG;
  if (n_win == 0) return AEW_OK;
  if (!d_sig || !d_sig_off || !d_peaks || !d_n_peaks || !d_fid || !d_features || !d_status) return AEW_ERR_ARG;
  const long long grid = n_win < 148LL * 16 ? n_win : 148LL * 16;
  AEW_LAUNCH(k6_features, (unsigned)grid, K6_THREADS, 0, stream, d_sig, d_sig_off, n_win, L, d_peaks, d_n_peaks,
             max_peaks, d_fid, d_valid, fs, d_features, d_status);
  return (int)cudaGetLastError();
}
