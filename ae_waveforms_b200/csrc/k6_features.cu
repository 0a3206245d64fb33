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

__device__ __forceinline__ double k6_value(int col, int b, const double (*X)[K6_MAXP], const double (*Y)[K6_MAXP]) {
  if (col < 20) return (col & 1) ? Y[col >> 1][b] : X[col >> 1][b];
  if (col < 30) {
    const int a = k6_dur[col - 20][0], c = k6_dur[col - 20][1];
    return __dsub_rn(X[c][b], X[a][b]);
  }
  const int pk = k6_prom[col - 30][0], b0 = k6_prom[col - 30][1], b1 = k6_prom[col - 30][2];
  double base = Y[b0][b];
  if (b1 >= 0) {                       // np.nanmax over the two baselines: NaN only if both are NaN
    const double o = Y[b1][b];
    if (isnan(base)) base = o;
    else if (!isnan(o) && o > base) base = o;
  }
  return __dsub_rn(Y[pk][b], base);
}

__global__ void __launch_bounds__(K6_THREADS)
k6_features(const double* __restrict__ sig, const long long* __restrict__ sig_off, long long n_win, int L,
            const int* __restrict__ peaks, const int* __restrict__ n_peaks, int max_peaks,
            const int* __restrict__ fid, const unsigned char* __restrict__ valid, double fs,
            double* __restrict__ out, int* __restrict__ status) {
  __shared__ double X[K6_NFID + 1][K6_MAXP];
  __shared__ double Y[K6_NFID + 1][K6_MAXP];
  __shared__ double vals[K6_THREADS / 32][K6_MAXP];
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
      if (f < K6_NFID) {
        const int idx = fw[f * max_peaks + b], r = pk[b];
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
      X[f][b] = x;
      Y[f][b] = y;
    }
    __syncthreads();
    for (int col = warp; col < K6_NOUT; col += K6_THREADS / 32) {
      // compact the non-NaN values of this column, beat order kept
      int m = 0;
      for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        const double v = b < nb ? k6_value(col, b, X, Y) : k6_nan();
        const bool ok = b < nb && !isnan(v);
        const unsigned mask = __ballot_sync(AEW_FULL, ok);
        if (ok) vals[warp][m + __popc(mask & ((1u << lane) - 1u))] = v;
        m += __popc(mask);
      }
      __syncwarp();
      if (m == 0) {
        if (lane == 0) row[col] = k6_nan();
        continue;
      }
      const int klo = (m - 1) >> 1, khi = m >> 1;
      for (int i = lane; i < m; i += 32) {
        const double v = vals[warp][i];
        int rank = 0;
        for (int j = 0; j < m; ++j) {
          const double u = vals[warp][j];
          rank += (u < v) || (u == v && j < i);
        }
        if (rank == klo) pick[warp][0] = v;
        if (rank == khi) pick[warp][1] = v;
      }
      __syncwarp();
      if (lane == 0) {
        const double lo = pick[warp][0], hi = pick[warp][1];
        row[col] = (klo == khi) ? lo : __ddiv_rn(__dadd_rn(lo, hi), 2.0);
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

}  // namespace

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
