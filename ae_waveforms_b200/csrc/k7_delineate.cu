// k7_delineate.cu -- K7: wavelet delineation of every beat of a 20 s window (P, Q, S, T peaks,
// P / QRS / T on- and offsets), the input of K6's np_features.
//
// Replaces nk.ecg_delineate(ecg_sample, r_peaks, sampling_rate)[1] as called at
// data_extract_vitaldb.py:131 (third party: neurokit2, un-vendored and version-unpinned in the
// reference; restated in oracle/nk_delineate.py from the published "dwt" method -- PARITY UNPINNED
// at that boundary, the kernel is checked against that restatement).
//
// One 512-thread block per window (two per SM: 64 registers, 88 KB shared memory), persistent, with a
// 1 MB slab of global scratch per block for the spline coefficients, the 2 kHz signal and two wavelet rows:
//   A  mean heart rate over the window: period between R peaks, monotone cubic (PCHIP) to every
//      sample, 60 / period, mean  -> epoch size for the Q / S search (ecg_segment)
//   B  warp per beat: Q = right-most prominent minimum left of R, S = left-most one from R on
//      (scipy find_peaks + peak_prominences semantics, prominence >= 5 % of the segment range)
//   C  cubic B-spline prefilter (mirror boundary; chunked recursion with a 32-sample run-in,
//      |pole|^32 = 5e-19) and evaluation at 2 kHz exactly as scipy.ndimage.zoom places its samples
//   D  a-trous quadratic-spline cascade: T_p = H_p T_{p-1} level by level, rows 2+d and 3+d kept
//      (d = int(log2(8 / (median rate / 60)))), including the library's left shift by 2^p that
//      leaves the last 2^p entries in place and the growing length of the intermediate
//   E  warp per beat: T / P peaks (zero crossing between significant +/- extrema, best score),
//      QRS bounds and P / T on- and offsets (threshold crossings next to the nearest wavelet
//      extremum), Python slice semantics for negative bounds, indices truncated back to fs
//   F  wrapper rules: a last entry >= L becomes NaN; an entry <= 0 marks the window "drop".
// Floating point: the index decisions are comparisons of float64 values that differ from the
// restatement's by rounding only (chunked prefilter, summation order), so indices can differ only
// where a comparison is within ~1e-15 of a tie.
#include "aew_rt.h"
#include "aew_pairwise.cuh"
#include <math.h>
#include <stdlib.h>

#ifndef K7_THREADS
#define K7_THREADS 512
#endif
#define K7_WARPS (K7_THREADS / 32)
#define K7_MAXP 72
#define K7_TILE 5120           // positions per shared-memory tile of the cascade (two buffers: 80 KB)
#define K7_RUNIN 32
#define K7_NANI (-0x7fffffffffffffffLL - 1)
#define K7_RATE 2000.0

namespace {

// debugging aid (AEW_K7_TIMING=1): thread 0 of every block adds the cycles since the previous mark to counter `slot`
#ifdef AEW_EMUL
#define K7_TICK(slot)
#define K7_CLOCK() 0
#else
#define K7_CLOCK() clock64()
#define K7_TICK(slot)                                                        \
  do {                                                                       \
    if (dbg && threadIdx.x == 0) {                                           \
      const long long t_now = clock64();                                     \
      atomicAdd(dbg + (slot), (unsigned long long)(t_now - t_last));         \
      t_last = t_now;                                                        \
    }                                                                        \
  } while (0)
#endif

struct K7Scratch {
  double *coef, *e2k, *wa, *wb;
};

#define MUL(a, b) __dmul_rn((a), (b))
#define ADD(a, b) __dadd_rn((a), (b))
#define SUB(a, b) __dsub_rn((a), (b))
#define DIV(a, b) __ddiv_rn((a), (b))

__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(AEW_FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(AEW_FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(AEW_FULL, v, o));
  return v;
}
__device__ __forceinline__ int hi_bit(unsigned m) { return 31 - __clz((int)m); }

// np.add.reduce order for n <= 128 contiguous float64
__device__ double np_sum_small(const double* a, int n) {
  if (n < 8) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i];
    return s;
  }
  double r[8];
  for (int k = 0; k < 8; ++k) r[k] = a[k];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int k = 0; k < 8; ++k) r[k] += a[i + k];
  double s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) s += a[i];
  return s;
}

// Python slice start:stop on a length-n sequence -> resolved start and length
__device__ __forceinline__ void py_slice(long long a, long long b, long long n, long long& a0, long long& len) {
  if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
  if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
  a0 = a;
  len = b > a ? b - a : 0;
}

__device__ __forceinline__ int sgn(double v) { return (v > 0.0) - (v < 0.0); }

// ------------------------------------------------------------------ B: Q / S peaks
struct Epoch {
  const double* sig;
  int L;
  int s0;
  __device__ __forceinline__ double at(int j) const {         // epoch sample j; 0 outside the signal (buffer)
    const int i = s0 + j;
    return (unsigned)i < (unsigned)L ? sig[i] : 0.0;
  }
};

// local maximum of x = -E over the segment [lo, lo + len) at position i (plateau start); returns the
// plateau's middle or -1 (scipy _local_maxima_1d)
__device__ __forceinline__ int seg_peak(const Epoch& e, int lo, int len, int i) {
  if (i < 1 || i > len - 2) return -1;
  const double x = -e.at(lo + i);
  if (!(-e.at(lo + i - 1) < x)) return -1;
  int ahead = i + 1;
  while (ahead < len - 1 && -e.at(lo + ahead) == x) ++ahead;
  if (!(-e.at(lo + ahead) < x)) return -1;
  return (i + ahead - 1) / 2;
}

// scipy peak_prominences (wlen = None) of the peak p of x = -E, evaluated by the whole warp: 32 samples per
// step in each direction up to the first one that is higher than the peak (or the segment's end)
__device__ double seg_prominence(const Epoch& e, int lo, int len, int p, int lane) {
  const double xp = -e.at(lo + p);
  double lmin = xp, rmin = xp;
  for (int pos = p - 1;; pos -= 32) {
    const int i = pos - lane;
    const double v = i >= 0 ? -e.at(lo + i) : 0.0;
    const unsigned stop = __ballot_sync(AEW_FULL, i < 0 || !(v <= xp));
    const int first = stop ? __ffs((int)stop) - 1 : 32;
    lmin = fmin(lmin, warp_min(lane < first ? v : INFINITY));
    if (stop) break;
  }
  for (int pos = p + 1;; pos += 32) {
    const int i = pos + lane;
    const double v = i < len ? -e.at(lo + i) : 0.0;
    const unsigned stop = __ballot_sync(AEW_FULL, i >= len || !(v <= xp));
    const int first = stop ? __ffs((int)stop) - 1 : 32;
    rmin = fmin(rmin, warp_min(lane < first ? v : INFINITY));
    if (stop) break;
  }
  return xp - fmax(lmin, rmin);
}

// first (which = 0) or last (which = 1) local minimum of the segment whose prominence is at least 5 % of
// the segment's range; -1 if none.  Executed by a whole warp.
__device__ int prominent_minimum(const Epoch& e, int lo, int len, int which, int lane) {
  if (len < 3) return -1;
  double mx = -INFINITY, mn = INFINITY;
  for (int i = lane; i < len; i += 32) {
    const double v = e.at(lo + i);
    mx = fmax(mx, v);
    mn = fmin(mn, v);
  }
  const double thr = 0.05 * (warp_max(mx) - warp_min(mn));
  const int n_chunks = (len + 31) / 32;
  for (int c = 0; c < n_chunks; ++c) {
    const int base = (which ? n_chunks - 1 - c : c) * 32;
    const int p = seg_peak(e, lo, len, base + lane);
    unsigned m = __ballot_sync(AEW_FULL, p >= 0);
    while (m) {                                    // candidates of this chunk in search order
      const int src = which ? hi_bit(m) : __ffs((int)m) - 1;
      m &= ~(1u << src);
      const int pp = __shfl_sync(AEW_FULL, p, src);
      if (!(seg_prominence(e, lo, len, pp, lane) < thr)) return pp;
    }
  }
  return -1;
}

// ------------------------------------------------------------------ E: searches on a wavelet row
struct Row {                    // a slice of a wavelet row, optionally negated
  const double* w;
  double sign;
  __device__ __forceinline__ double at(int i) const { return sign * w[i]; }
};

// local maximum of r over [a0, a0 + len) at plateau start i -> plateau middle or -1
__device__ __forceinline__ int row_peak(const Row& r, int len, int i, bool use_abs) {
  if (i < 1 || i > len - 2) return -1;
  double x = r.at(i), l = r.at(i - 1);
  if (use_abs) { x = fabs(x); l = fabs(l); }
  if (!(l < x)) return -1;
  int ahead = i + 1;
  while (ahead < len - 1 && (use_abs ? fabs(r.at(ahead)) : r.at(ahead)) == x) ++ahead;
  const double nx = use_abs ? fabs(r.at(ahead)) : r.at(ahead);
  if (!(nx < x)) return -1;
  return (i + ahead - 1) / 2;
}

// _dwt_delineate_tp_peaks for one beat: zero crossing between a significant positive and the following
// significant negative extremum with the best score.  mode 0: T (score = y - (t - ref)), 1: P (y - |t - ref|).
__device__ long long tp_peak(const double* row, const double* e2k, long long n2k, long long a, long long b,
                             double height_w, int mode, double ref, int lane) {
  long long a0, len64;
  py_slice(a, b, n2k, a0, len64);
  if (len64 == 0) return K7_NANI;
  const int len = (int)len64;
  const Row r{row + a0, 1.0};
  const double* ecg = e2k + a0;
  double ss = 0.0, mx = -INFINITY;
  {
    double ss2 = 0.0, mx2 = -INFINITY;
    int i = lane;
    for (; i + 32 < len; i += 64) {
      const double v = r.at(i), u = r.at(i + 32);
      ss += v * v;
      ss2 += u * u;
      mx = fmax(mx, v);
      mx2 = fmax(mx2, u);
    }
    if (i < len) {
      const double v = r.at(i);
      ss += v * v;
      mx = fmax(mx, v);
    }
    ss += ss2;
    mx = fmax(mx, mx2);
  }
  const double height = height_w * sqrt(warp_sum(ss) / (double)len);
  const double floor_abs = 0.025 * warp_max(mx);
  int carry = r.at(0) > 0.0 ? 0 : -1;
  double best = -INFINITY;
  int best_z = -1;
  for (int base = 0; base < len; base += 32) {
    int p = row_peak(r, len, base + lane, true);
    if (p >= 0) {
      const double h = fabs(r.at(p));
      if (!(h >= height) || !(h > floor_abs)) p = -1;
    }
    const unsigned m = __ballot_sync(AEW_FULL, p >= 0);
    const unsigned lower = m & ((1u << lane) - 1u);
    const int from_lane = __shfl_sync(AEW_FULL, p, lower ? hi_bit(lower) : 0);
    const int prev = lower ? from_lane : carry;
    if (p >= 0 && prev >= 0 && r.at(prev) > 0.0 && r.at(p) < 0.0) {
      int z = prev;
      int sz = sgn(r.at(z));
      while (z < p) {                              // first sample whose successor has another sign
        const int sn = sgn(r.at(z + 1));
        if (sn != sz) break;
        ++z;
      }
      const double t = (double)z / K7_RATE;
      const double y = ecg[z];
      const double sc = mode == 0 ? y - (t - ref) : y - fabs(t - ref);
      if (sc > best) { best = sc; best_z = z; }          // this lane's candidates come in increasing z
    }
    const int last = __shfl_sync(AEW_FULL, p, m ? hi_bit(m) : 0);
    if (m) carry = last;
  }
  // np.argmax: first maximum = the smallest z among equal scores
  for (int o = 16; o; o >>= 1) {
    const double ob = __shfl_xor_sync(AEW_FULL, best, o);
    const int oz = __shfl_xor_sync(AEW_FULL, best_z, o);
    if (oz >= 0 && (best_z < 0 || ob > best || (ob == best && oz < best_z))) { best = ob; best_z = oz; }
  }
  return best_z < 0 ? K7_NANI : (long long)best_z + a;
}

// last sample before the last local maximum of sign*row[a:b] that is below weight * that maximum
__device__ long long onset_search(const double* row, long long n2k, long long a, long long b, double sign,
                                  double weight, int lane) {
  if (a == K7_NANI || b == K7_NANI) return K7_NANI;
  long long a0, len64;
  py_slice(a, b, n2k, a0, len64);
  if (len64 < 3) return K7_NANI;
  const int len = (int)len64;
  const Row r{row + a0, sign};
  int pk = -1;
  for (int base = ((len - 1) / 32) * 32; base >= 0 && pk < 0; base -= 32) {
    const int p = row_peak(r, len, base + lane, false);
    const unsigned m = __ballot_sync(AEW_FULL, p >= 0);
    if (m) pk = __shfl_sync(AEW_FULL, p, hi_bit(m));
  }
  if (pk < 0) return K7_NANI;
  const double eps = weight * r.at(pk);
  for (int base = ((pk - 1) / 32) * 32; pk > 0 && base >= 0; base -= 32) {
    const int i = base + lane;
    const unsigned m = __ballot_sync(AEW_FULL, i < pk && r.at(i) < eps);
    if (m) return (long long)(base + hi_bit(m)) + a;
  }
  return K7_NANI;
}

// first sample from the first local maximum of sign*row[a:b] on that is below weight * that maximum
__device__ long long offset_search(const double* row, long long n2k, long long a, long long b, double sign,
                                   double weight, int lane) {
  if (a == K7_NANI || b == K7_NANI) return K7_NANI;
  long long a0, len64;
  py_slice(a, b, n2k, a0, len64);
  if (len64 < 3) return K7_NANI;
  const int len = (int)len64;
  const Row r{row + a0, sign};
  int pk = -1;
  for (int base = 0; base < len && pk < 0; base += 32) {
    const int p = row_peak(r, len, base + lane, false);
    const unsigned m = __ballot_sync(AEW_FULL, p >= 0);
    if (m) pk = __shfl_sync(AEW_FULL, p, __ffs((int)m) - 1);
  }
  if (pk < 0) return K7_NANI;
  const double eps = weight * r.at(pk);
  for (int base = (pk / 32) * 32; base < len; base += 32) {
    const int i = base + lane;
    const unsigned m = __ballot_sync(AEW_FULL, i >= pk && i < len && r.at(i) < eps);
    if (m) return (long long)(base + (__ffs((int)m) - 1)) + a;
  }
  return K7_NANI;
}

// _dwt_resample_points(x, 2000, fs): int(x * fs / 2000), truncation toward zero
__device__ __forceinline__ int back_to_fs(long long v, double fs) {
  if (v == K7_NANI) return AEW_FID_NAN;
  return (int)(((double)v * fs) / K7_RATE);
}

__device__ __forceinline__ double k7_round3(double v) { return rint(v * 1000.0) / 1000.0; }   // np.round(v, 3)

// int(np.log2(x)) for x > 0
__device__ __forceinline__ int int_log2(double x) {
  int e;
  const double m = frexp(x, &e);          // x = m * 2^e, m in [0.5, 1)
  const int fl = e - 1;                   // floor(log2 x)
  if (fl >= 0) return fl;
  return m == 0.5 ? fl : fl + 1;          // truncation toward zero for negative logarithms
}

// 60 / period at sample i of the window: scipy PPoly evaluation of the PCHIP through (peak, period), constant
// outside the first / last peak (signal_interpolate's out-of-bounds rule), NumPy's operation order
struct RateAt {
  const int* pk;
  const double (*c)[K7_MAXP];
  int nb;
  __device__ __forceinline__ double operator()(long long i, bool) const {
    const int p_first = pk[0], p_last = pk[nb - 1];
    const int xe = i < p_first ? p_first : (i > p_last ? p_last : (int)i);
    int lo = 0, hi = nb - 2;                     // interval k with pk[k] <= xe < pk[k+1] (last knot: nb - 2)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (pk[mid] <= xe) lo = mid; else hi = mid - 1;
    }
    const double s = (double)(xe - pk[lo]);
    double res = c[3][lo], z = s;
    res = ADD(res, MUL(c[2][lo], z));
    z = MUL(z, s);
    res = ADD(res, MUL(c[1][lo], z));
    z = MUL(z, s);
    res = ADD(res, MUL(c[0][lo], z));
    return DIV(60.0, res);
  }
};

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(K7_THREADS, 2)
k7_delineate(const double* __restrict__ sig_all, const long long* __restrict__ sig_off, long long n_win, int L,
             long long n2k, double fs, const int* __restrict__ peaks, const int* __restrict__ n_peaks, int max_peaks,
             const unsigned char* __restrict__ select, int* __restrict__ fid, unsigned char* __restrict__ flags,
             double* __restrict__ scratch, long long scratch_stride, unsigned long long* __restrict__ dbg) {
  __shared__ int s_pk[K7_MAXP];
  __shared__ long long s_r2k[K7_MAXP];
  __shared__ double s_y[K7_MAXP], s_d[K7_MAXP], s_c[4][K7_MAXP];
  __shared__ double s_red[K7_WARPS];
  __shared__ double s_vals[128];               // leaf slots of np.mean's pairwise tree (L <= 16,384)
  __shared__ unsigned char s_valid[128];
  __shared__ long long s_tp[2][K7_MAXP];        // T and P peaks at 2 kHz
  __shared__ double s_par[8];
  __shared__ int s_ipar[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  K7Scratch S;
  {
    double* base = scratch + 16 + (long long)blockIdx.x * scratch_stride;   // 128 bytes: the window counter
    S.coef = base;
    S.e2k = S.coef + ((L + 7) & ~7);
    S.wa = S.e2k + ((n2k + 7) & ~7LL);
    S.wb = S.wa + ((n2k + 7) & ~7LL);
  }
  const double zp = sqrt(3.0) - 2.0;             // pole of the cubic B-spline prefilter
  AEW_DYN_SMEM(dyn);
  double* const sm = reinterpret_cast<double*>(dyn);   // 2 * K7_TILE doubles: the window, then the cascade tiles

  // Windows are CLAIMED, not strided over: the screen drops ~16 % of them and the cost of a kept one grows with its
  // beat count, so a fixed share per block leaves the slowest block ~17 % behind the mean at 60 windows per block.
  unsigned long long* const next_window = reinterpret_cast<unsigned long long*>(scratch);
  __shared__ long long s_w;
  for (;;) {
    __syncthreads();                              // previous window's shared state (and s_w) is no longer read
    if (tid == 0) s_w = (long long)atomicAdd(next_window, 1ull);
    __syncthreads();
    const long long w = s_w;
    if (w >= n_win) break;
    if (select && !select[w]) continue;
    long long t_last = dbg ? K7_CLOCK() : 0;
    (void)t_last;
    const double* sig = sig_all + sig_off[w];
    int nb = n_peaks[w];
    if (nb > max_peaks) nb = max_peaks;
    int* fw = fid + w * 10LL * max_peaks;
    __syncthreads();                              // previous window's shared state is no longer read
    for (int i = tid; i < 10 * max_peaks; i += K7_THREADS) fw[i] = AEW_FID_NAN;
    // signal_period: <= 3 peaks -> NaN period -> exception in the library -> NaN features, window kept
    if (nb <= 3 || (double)L < fs * 4.0) {
      if (tid == 0) flags[w] = 1;
      continue;
    }
    for (int i = tid; i < L; i += K7_THREADS) sm[i] = sig[i];   // the window stays in shared memory for A - C
    for (int i = tid; i < nb; i += K7_THREADS) {
      s_pk[i] = peaks[w * (long long)max_peaks + i];
      s_r2k[i] = (long long)((double)((long long)s_pk[i] * 2000LL) / fs);
    }
    __syncthreads();

    // ---- A: mean heart rate (signal_rate with desired_length = L) and median rate at 2 kHz.
    // Bit-identical to NumPy / SciPy (explicit round-to-nearest operations in their order, np.mean's
    // pairwise tree): the epoch labels below hit 0.0 EXACTLY whenever 0.35 (n - 1) is an integer, so the
    // slice decisions depend on the last bit of this mean.
    if (tid >= 1 && tid < nb) s_y[tid] = DIV((double)(s_pk[tid] - s_pk[tid - 1]), fs);
    __syncthreads();
    if (tid == 0) s_y[0] = DIV(np_sum_small(s_y + 1, nb - 1), (double)(nb - 1));
    __syncthreads();
    if (tid < nb) {                               // PCHIP end-point and interior derivatives (scipy _find_derivatives)
      const int i = tid;
      double d;
      if (i == 0 || i == nb - 1) {
        const int k0 = i == 0 ? 0 : nb - 2, k1 = i == 0 ? 1 : nb - 3;
        const double h0 = (double)(s_pk[k0 + 1] - s_pk[k0]), h1 = (double)(s_pk[k1 + 1] - s_pk[k1]);
        const double m0 = DIV(SUB(s_y[k0 + 1], s_y[k0]), h0), m1 = DIV(SUB(s_y[k1 + 1], s_y[k1]), h1);
        d = DIV(SUB(MUL(ADD(MUL(2.0, h0), h1), m0), MUL(h0, m1)), ADD(h0, h1));
        if (sgn(d) != sgn(m0)) d = 0.0;
        else if (sgn(m0) != sgn(m1) && fabs(d) > MUL(3.0, fabs(m0))) d = MUL(3.0, m0);
      } else {
        const double h0 = (double)(s_pk[i] - s_pk[i - 1]), h1 = (double)(s_pk[i + 1] - s_pk[i]);
        const double m0 = DIV(SUB(s_y[i], s_y[i - 1]), h0), m1 = DIV(SUB(s_y[i + 1], s_y[i]), h1);
        if (sgn(m0) != sgn(m1) || m0 == 0.0 || m1 == 0.0) d = 0.0;
        else {
          const double w1 = ADD(MUL(2.0, h1), h0), w2 = ADD(h1, MUL(2.0, h0));
          d = DIV(1.0, DIV(ADD(DIV(w1, m0), DIV(w2, m1)), ADD(w1, w2)));
        }
      }
      s_d[i] = d;
    }
    __syncthreads();
    if (tid < nb - 1) {                           // CubicHermiteSpline coefficients, power basis at the left knot
      const double dx = (double)(s_pk[tid + 1] - s_pk[tid]);
      const double slope = DIV(SUB(s_y[tid + 1], s_y[tid]), dx);
      const double t = DIV(SUB(ADD(s_d[tid], s_d[tid + 1]), MUL(2.0, slope)), dx);
      s_c[0][tid] = DIV(t, dx);
      s_c[1][tid] = SUB(DIV(SUB(slope, s_d[tid]), dx), t);
      s_c[2][tid] = s_d[tid];
      s_c[3][tid] = s_y[tid];
    }
    __syncthreads();
    {
      RateAt f{s_pk, s_c, nb};
      const double tot = pw_block_sum(f, L, pw_depth(L), s_vals, s_valid);   // np.mean's summation order
      if (tid == 0) s_red[0] = tot;
    }
    __syncthreads();
    if (warp == 0) {
      // median of 60 / period at 2 kHz (np.median; first period = mean of the others);
      // s_d becomes the period list (the PCHIP derivatives are consumed)
      for (int i = lane; i < nb; i += 32)
        if (i > 0) s_d[i] = DIV((double)(s_r2k[i] - s_r2k[i - 1]), K7_RATE);
      __syncwarp();
      if (lane == 0) s_d[0] = DIV(np_sum_small(s_d + 1, nb - 1), (double)(nb - 1));
      __syncwarp();
      double lo_v = 0.0, hi_v = 0.0;
      const int klo = (nb - 1) >> 1, khi = nb >> 1;
      for (int i = lane; i < nb; i += 32) {
        const double v = DIV(60.0, s_d[i]);
        int rank = 0;
        for (int j = 0; j < nb; ++j) {
          const double u = DIV(60.0, s_d[j]);
          rank += (u < v) || (u == v && j < i);
        }
        if (rank == klo) lo_v = v;
        if (rank == khi) hi_v = v;
      }
      lo_v = warp_sum(lo_v);                       // exactly one lane holds each (others add 0)
      hi_v = warp_sum(hi_v);
      if (lane == 0) {
        const double heart_rate = DIV(s_red[0], (double)L);
        const double window_size = DIV(60.0, heart_rate);
        s_par[0] = -MUL(0.35, window_size);        // epochs_start
        s_par[1] = MUL(SUB(1.0, 0.35), window_size);   // epochs_end
        const double avg = klo == khi ? lo_v : DIV(ADD(lo_v, hi_v), 2.0);
        const double beat = DIV(60.0, avg);
        const int deg = int_log2(DIV(DIV(K7_RATE, 250.0), DIV(avg, 60.0)));
        s_ipar[0] = deg;
        s_par[2] = k7_round3(MUL(0.2, beat));      // p2r_duration
        s_par[3] = k7_round3(MUL(0.25, beat));     // rt_duration
        s_par[4] = k7_round3(MUL(0.3, beat));      // duration_onset / duration_offset
        s_par[5] = k7_round3(MUL(0.6, beat));      // T duration_onset
        s_ipar[1] = (isnan(avg) || isnan(heart_rate) || deg < 0 || deg > 5) ? 1 : 0;   // rows 2+d, 3+d of 9
      }
    }
    __syncthreads();
    if (s_ipar[1]) {
      if (tid == 0) flags[w] = 1;
      continue;
    }

    K7_TICK(0);
    // ---- B: Q and S peaks (warp per beat) on the signal at fs
    {
      const double start_s = s_par[0], end_s = s_par[1];
      const long long buf = (long long)MUL(SUB(end_s, start_s), fs);
      int bad = 0;
      for (int b = warp; b < nb; b += K7_WARPS) {
        const long long r = s_pk[b];
        const long long ea = (long long)ADD((double)(r + buf), MUL(start_s, fs));
        const long long eb = (long long)ADD((double)(r + buf), MUL(end_s, fs));
        const long long n = eb - ea;
        long long n_left = 0, right0 = 0;
        if (n >= 2) {                              // labels = np.linspace(start_s, end_s, n)
          const double step = DIV(SUB(end_s, start_s), (double)(n - 1));
          long long j = (long long)floor(-start_s / step);
          if (j > n - 2) j = n - 2;
          if (j < 0) j = 0;
          while (j + 1 <= n - 2 && ADD(MUL((double)(j + 1), step), start_s) <= 0.0) ++j;
          while (j >= 0 && ADD(MUL((double)j, step), start_s) > 0.0) --j;
          n_left = j + 1;                          // the last label is end_s > 0
          right0 = n_left - ((j >= 0 && ADD(MUL((double)j, step), start_s) == 0.0) ? 1 : 0);
        } else {
          n_left = n;                              // no label above 0 -> np.min([]) raises
        }
        if (n_left >= n) { bad = 1; continue; }
        const Epoch e{sm, L, (int)(ea - buf)};
        const int qi = prominent_minimum(e, 0, (int)n_left, 1, lane);
        long long n_right = n - right0;
        if (b == nb - 1 && e.s0 + n > L) n_right = (long long)L - e.s0 - right0;   // last beat: NaN past the end
        const int si = n_right > 0 ? prominent_minimum(e, (int)right0, (int)n_right, 0, lane) : -1;
        if (lane == 0) {
          fw[3 * max_peaks + b] = qi < 0 ? AEW_FID_NAN : (int)(r - (n_left - qi));
          fw[6 * max_peaks + b] = si < 0 ? AEW_FID_NAN : (int)(r + si);
        }
      }
      if (__syncthreads_or(bad)) {
        if (tid == 0) flags[w] = 1;
        continue;
      }
    }

    K7_TICK(1);
    // ---- C: cubic B-spline coefficients of the window (scipy.ndimage.zoom, order 3, mirror prefilter), in place
    // in shared memory: every thread owns a chunk; the state entering a chunk comes from a 32-sample run-in
    // over its neighbour's samples (read before anybody writes), then the chunk is overwritten.
    {
      const int ch = (L + K7_THREADS - 1) / K7_THREADS;
      const int c0 = tid * ch, c1 = c0 + ch < L ? c0 + ch : L;
      double c = 0.0;
      if (c0 < c1)
        for (int i = c0 - K7_RUNIN; i < c0; ++i) {
          int k = i < 0 ? -i : i;
          if (k >= L) k = L - 1;
          c = fma(zp, c, 6.0 * sm[k]);
        }
      __syncthreads();
      for (int i = c0; i < c1; ++i) {
        c = fma(zp, c, 6.0 * sm[i]);
        sm[i] = c;
      }
      __syncthreads();
      // anticausal pass: c[i] = zp (c[i+1] - c+[i]), c[L-1] = zp / (zp^2 - 1) (c+[L-1] + zp c+[L-2])
      if (c0 < c1) {
        int i;
        if (c1 + K7_RUNIN >= L) {
          c = (zp / (zp * zp - 1.0)) * (sm[L - 1] + zp * sm[L - 2]);
          i = L - 1;
        } else {
          c = 0.0;
          i = c1 + K7_RUNIN;
        }
        for (--i; i >= c1 - 1 && i >= c0; --i) c = zp * (c - sm[i]);   // c = coefficient at c1 - 1 (or L - 1 itself)
      }
      __syncthreads();
      if (c0 < c1) {
        // here c is the coefficient of position c1 - 1
        sm[c1 - 1] = c;
        for (int i = c1 - 2; i >= c0; --i) {
          c = zp * (c - sm[i]);
          sm[i] = c;
        }
      }
      __syncthreads();
      for (int i = tid; i < L; i += K7_THREADS) S.coef[i] = sm[i];
      __syncthreads();
    }

    K7_TICK(2);
    // ---- D: 2 kHz samples and the a-trous cascade, tile by tile in shared memory.  A tile holds K7_TILE
    // consecutive positions of every level; level p is valid 2 * 2^p less far to the left and 2^p less far to
    // the right than level p - 1, so the halo is sized for the rows qa = 2 + d and qb = 3 + d that are kept.
    {
      double* bufs[2] = {sm, sm + K7_TILE};
      const int deg = s_ipar[0];
      const int qa = 2 + deg, qb = 3 + deg;
      const int halo_l = 2 * ((1 << (qa + 1)) - 1), halo_r = (1 << (qa + 2)) - 1;
      const int ts = K7_TILE - halo_l - halo_r;
      const double ratio = (double)(L - 1) / (double)(n2k - 1);
      const long long dqa = 1LL << qa, dqb = 1LL << qb;
      for (long long t0 = 0; t0 < n2k; t0 += ts) {
        const long long base = t0 - halo_l;
        const long long t1 = t0 + ts < n2k ? t0 + ts : n2k;
        double* prev = bufs[0];
        double* cur = bufs[1];
        // the spline coefficients this tile touches, mirrored at the ends, staged in the other buffer
        const long long j_lo = base > 0 ? base : 0;
        const long long j_hi = (base + K7_TILE < n2k ? base + K7_TILE : n2k) - 1;
        const int i_lo = (int)floor((double)j_lo * ratio) - 1;
        const int i_cnt = (int)floor((double)j_hi * ratio) + 3 - i_lo;
        for (int k = tid; k < i_cnt; k += K7_THREADS) {
          int idx = i_lo + k;
          if (idx < 0) idx = -idx;
          if (idx >= L) idx = 2 * L - 2 - idx;
          cur[k] = S.coef[idx];
        }
        __syncthreads();
        for (int i = tid; i < K7_TILE; i += K7_THREADS) {   // level -1: the resampled signal
          const long long j = base + i;
          double t = 0.0;
          if (j >= 0 && j < n2k) {
            const double cc = (double)j * ratio;
            const double fl = floor(cc);
            const double y = cc - fl, z = 1.0 - y;
            const double w1 = (y * y * (y - 2.0) * 3.0 + 4.0) * (1.0 / 6.0);
            const double w2 = (z * z * (z - 2.0) * 3.0 + 4.0) * (1.0 / 6.0);
            const double w0 = z * z * z * (1.0 / 6.0);
            const double w3 = 1.0 - w0 - w1 - w2;
            const double* cs = cur + ((int)fl - 1 - i_lo);
            t = cs[0] * w0;
            t += cs[1] * w1;
            t += cs[2] * w2;
            t += cs[3] * w3;
            if (j >= t0 && j < t1) S.e2k[j] = t;
          }
          prev[i] = t;
        }
        __syncthreads();
        K7_TICK(3);
        long long len_prev = n2k;                  // true length of the previous level (it grows by 3 * 2^p)
        for (int p = 0; p <= qa; ++p) {
          const int d = 1 << p;
          if (p == qa) {                           // row qa = G_qa(T_{qa-1})
            for (long long n = t0 + tid; n < t1; n += K7_THREADS) {
              const double ahead = n + dqa < len_prev ? prev[n + dqa - base] : 0.0;
              S.wa[n] = 2.0 * ahead - 2.0 * prev[n - base];
            }
          }
          const long long len_cur = len_prev + 3 * d;
          const long long shifted = len_cur - d;
          // interior positions [f_lo, f_hi): every tap lies inside the signal, the buffer and the shifted part
          long long lo64 = 2LL * d - base, hi64 = n2k - d - base;
          const int f_lo = lo64 > 2 * d ? (int)(lo64 < K7_TILE ? lo64 : K7_TILE) : 2 * d;
          const int f_hi = hi64 < K7_TILE - d ? (int)(hi64 > 0 ? hi64 : 0) : K7_TILE - d;
#pragma unroll 4
          for (int i = tid; i < K7_TILE; i += K7_THREADS) {
            double acc = 0.0;
            if (i >= f_lo && i < f_hi) {
              acc = 0.125 * prev[i + d];
              acc += 0.375 * prev[i];
              acc += 0.375 * prev[i - d];
              acc += 0.125 * prev[i - 2 * d];
            } else {
              const long long n = base + i;
              if (n >= 0 && n < len_cur) {
                const long long m = n < shifted ? n + d : n;
                const int im = (int)(m - base);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int ix = im - k * d;
                  const long long x = m - (long long)k * d;
                  const double v = (ix >= 0 && ix < K7_TILE && x >= 0 && x < len_prev) ? prev[ix] : 0.0;
                  acc += ((k == 0 || k == 3) ? 0.125 : 0.375) * v;
                }
              }
            }
            cur[i] = acc;
          }
          __syncthreads();
          double* sw = prev; prev = cur; cur = sw;
          len_prev = len_cur;
        }
        for (long long n = t0 + tid; n < t1; n += K7_THREADS) {    // row qb = G_qb(T_qa)
          const double ahead = n + dqb < len_prev ? prev[n + dqb - base] : 0.0;
          S.wb[n] = 2.0 * ahead - 2.0 * prev[n - base];
        }
        __syncthreads();
        K7_TICK(4);
      }
    }

    // ---- E: per-beat searches at 2 kHz, a warp per task: first the T and P peaks, then the six bounds
    {
      const double p2r = s_par[2], rt = s_par[3], d_on = s_par[4], d_on_t = s_par[5];
      const long long bndry = (long long)(0.5 * 0.13 * K7_RATE);
      for (int task = warp; task < 2 * nb; task += K7_WARPS) {
        const int b = task >> 1;
        const long long r = s_r2k[b];
        long long v;
        if ((task & 1) == 0)
          v = tp_peak(S.wb, S.e2k, n2k, r + bndry, r + 2 * (long long)(rt * K7_RATE), 0.25, 0, rt - 0.5 * 0.13, lane);
        else
          v = tp_peak(S.wa, S.e2k, n2k, r - 2 * (long long)(p2r * K7_RATE), r - bndry, 0.02, 1, p2r, lane);
        if (lane == 0) {
          s_tp[task & 1][b] = v;
          fw[((task & 1) ? 0 : 7) * max_peaks + b] = back_to_fs(v, fs);
        }
      }
      __syncthreads();
      K7_TICK(5);
      const long long n_on = (long long)(d_on * K7_RATE), n_on_t = (long long)(d_on_t * K7_RATE);
      for (int task = warp; task < 6 * nb; task += K7_WARPS) {
        const int b = task / 6, kind = task - 6 * b;
        const long long r = s_r2k[b], t = s_tp[0][b], p = s_tp[1][b];
        long long v;
        int slot;
        if (kind == 0) {                           // QRS onset: between the P peak and the Q peak
          const int q500 = fw[3 * max_peaks + b];
          const long long q2k = q500 == AEW_FID_NAN ? K7_NANI : (long long)((double)((long long)q500 * 2000LL) / fs);
          v = onset_search(S.wa, n2k, p, q2k, -1.0, 0.5, lane);
          slot = 4;
        } else if (kind == 1) {                    // QRS offset: between the R peak and the T peak
          v = offset_search(S.wa, n2k, r, t, 1.0, 0.5, lane);
          slot = 5;
        } else if (kind == 2) {
          v = onset_search(S.wa, n2k, p == K7_NANI ? K7_NANI : p - n_on, p, 1.0, 0.4, lane);
          slot = 1;
        } else if (kind == 3) {
          v = offset_search(S.wa, n2k, p, p == K7_NANI ? K7_NANI : p + n_on, -1.0, 0.4, lane);
          slot = 2;
        } else if (kind == 4) {
          v = onset_search(S.wa, n2k, t == K7_NANI ? K7_NANI : t - n_on_t, t, 1.0, 0.6, lane);
          slot = 8;
        } else {
          v = offset_search(S.wa, n2k, t, t == K7_NANI ? K7_NANI : t + n_on, -1.0, 0.4, lane);
          slot = 9;
        }
        if (lane == 0) fw[slot * max_peaks + b] = back_to_fs(v, fs);
      }
    }
    __syncthreads();
    K7_TICK(6);

    // ---- F: wrapper rules of ecg_delineate
    {
      int drop = 0;
      if (tid < 10) {
        int* lst = fw + tid * max_peaks;
        if (lst[nb - 1] != AEW_FID_NAN && lst[nb - 1] >= L) lst[nb - 1] = AEW_FID_NAN;
        for (int b = 0; b < nb; ++b)
          if (lst[b] != AEW_FID_NAN && lst[b] <= 0) drop = 1;
      }
      drop = __syncthreads_or(drop);
      if (tid == 0) flags[w] = drop ? 2 : 0;
    }
    K7_TICK(7);
  }
}

}  // namespace

static unsigned long long* g_k7_dbg = nullptr;   // device counters, allocated once when AEW_K7_TIMING is set
static unsigned long long* k7_timing_buffer() {
#ifndef AEW_EMUL
  if (!g_k7_dbg && getenv("AEW_K7_TIMING")) {
    if (cudaMalloc(&g_k7_dbg, 8 * sizeof(unsigned long long)) != cudaSuccess) g_k7_dbg = nullptr;
    else cudaMemset(g_k7_dbg, 0, 8 * sizeof(unsigned long long));
  }
#endif
  return g_k7_dbg;
}

// debugging aid (AEW_K7_TIMING=1): cycles thread 0 of every block spent per phase of k7_delineate, summed over
// blocks and windows: A heart rate, B Q/S peaks, C prefilter, D coefficient staging + resampling, D cascade + rows,
// E T/P peaks, E bounds, F wrapper rules.  Copies and clears the 8 counters (synchronises the device).
extern "C" int aew_delineate_timing_dump(unsigned long long* h_out8) {
#ifdef AEW_EMUL
  for (int i = 0; i < 8; ++i) h_out8[i] = 0;
  return AEW_OK;
#else
  if (!g_k7_dbg) {
    for (int i = 0; i < 8; ++i) h_out8[i] = 0;
    return AEW_OK;
  }
  cudaError_t e = cudaMemcpy(h_out8, g_k7_dbg, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemset(g_k7_dbg, 0, 8 * sizeof(unsigned long long));
  return (int)e;
#endif
}

extern "C" int aew_delineate_blocks(long long n_win) {
  const long long cap = 148LL * 2;
  return (int)(n_win < cap ? (n_win > 0 ? n_win : 1) : cap);
}

static long long k7_n2k(int L, double fs) { return (long long)rint((double)L * K7_RATE / fs); }

static long long k7_stride(int L, long long n2k) { return ((L + 7) & ~7) + 3 * ((n2k + 7) & ~7LL); }

extern "C" size_t aew_delineate_workspace_bytes(long long n_win, int L, double fs) {
  if (L <= 0 || !(fs > 0)) return 0;
  return (size_t)aew_delineate_blocks(n_win) * (size_t)k7_stride(L, k7_n2k(L, fs)) * sizeof(double) + 512;
}

extern "C" int aew_delineate_f64(const double* d_sig, const long long* d_sig_off, long long n_win, int L, double fs,
                                 const int* d_peaks, const int* d_n_peaks, int max_peaks,
                                 const unsigned char* d_select, int* d_fid, unsigned char* d_flags, void* d_ws,
                                 size_t ws_bytes, void* stream) {
  if (n_win < 0 || L < 64 || L > 2 * K7_TILE || max_peaks <= 0 || max_peaks > K7_MAXP || !(fs > 0)) return AEW_ERR_ARG;
  if (n_win == 0) return AEW_OK;
  if (!d_sig || !d_sig_off || !d_peaks || !d_n_peaks || !d_fid || !d_flags || !d_ws) return AEW_ERR_ARG;
  if (ws_bytes < aew_delineate_workspace_bytes(n_win, L, fs)) return AEW_ERR_WORKSPACE;
  const long long n2k = k7_n2k(L, fs);
  // the analysis rate must be at least twice the input rate (a tile stages K7_TILE * L / n2k + 6 coefficients)
  if (n2k < 2LL * L || n2k > (1LL << 30)) return AEW_ERR_UNSUPPORTED;
  double* ws = (double*)(((uintptr_t)d_ws + 127) & ~(uintptr_t)127);
  const size_t smem = 2 * K7_TILE * sizeof(double);
  if (AEW_SET_MAX_SMEM(k7_delineate, smem) != 0) return AEW_ERR_UNSUPPORTED;
  cudaMemsetAsync(ws, 0, 128, (cudaStream_t)stream);                  // the kernel's window counter
  AEW_LAUNCH(k7_delineate, (unsigned)aew_delineate_blocks(n_win), K7_THREADS, smem, stream, d_sig, d_sig_off, n_win, L,
             n2k, fs, d_peaks, d_n_peaks, max_peaks, d_select, d_fid, d_flags, ws, k7_stride(L, n2k),
             k7_timing_buffer());
  return (int)cudaGetLastError();
}
