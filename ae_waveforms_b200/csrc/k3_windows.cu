// k3_windows.cu -- per-window z-score, R-peak detection, physiological screen
// and median beat (K3 + K4, fused: the window is read from HBM once).
//
// One 512-thread block per 20 s window (L = 10,000 samples, 80 KB, staged once in
// shared memory).  Replaces, per window,
//   data_extract_vitaldb.py:93          per-window zscore            (bit-exact)
//   remove_samples.py:31-33,43          drop windows containing NaN
//   remove_samples.py:50-51             nk.ecg_findpeaks (default "neurokit" method,
//                                       SURVEY.md appendix B.2 / oracle/nk_peaks.py)
//   remove_samples.py:53-64,85-121      bounds filter, estimate_hr, rmssd, np.std, sd/n
//   remove_samples.py:70-74             keep/remove mask                (bit-exact)
//   median_sample.py:6-54               320-sample R-aligned sections, column-wise
//                                       nanmedian -> np_waves row       (bit-exact)
//
// Exactness.  Everything downstream of the QRS indicator works on integers or on
// NumPy-ordered reductions (aew_pairwise.cuh) and is bit-identical to the
// reference.  The indicator smoothgrad > 1.5*avggrad comes from two
// scipy.ndimage.uniform_filter1d calls whose C code is ONE running sum over the
// window (appendix B.3): 2 x 10,000 dependent additions.  The kernel evaluates both
// box filters with block-wide prefix sums instead (different rounding, ~1e-13),
// flags a window as a HAZARD when any sample's comparison is closer to a tie than a
// bound that dominates both evaluations' rounding error, and only for those
// (about 1 window in 10^5) thread 0 re-runs SciPy's literal running sums.  Integer
// outputs are therefore identical to the sequential evaluation for every window.
// mode 1 forces the sequential evaluation everywhere (parity suite); mode 2
// disables the fallback (measurement only).
//
// Round-1 profile (profiles/r1_k3_before.txt) showed the kernel instruction-bound,
// 48 % of the instructions in a branchy warp-shuffle median.  The median is now one
// thread per beat column with a 32-input Batcher network held in registers
// (191 compare-exchanges of DSETP + 4 FSEL); more than 32 beats per window fall back
// to a branch-free warp-bitonic sort.  Leaf ranges of the NumPy tree are tabulated
// once per block; HR / RMSSD are computed cooperatively instead of by thread 0.
//
// Algorithmic HBM bytes: 8 B per input sample read + (320*12 + 72*4 + 48) B per
// window written.  Bound: instruction issue / FP64, not HBM.
#include "aew_pairwise.cuh"
#include "aew_math.cuh"
#include "../../include/aewave.h"

#define K3_NT 512
#define K3_ITEMS 21                       // odd -> conflict-free per-thread strips
#define K3_MAXL (K3_NT * K3_ITEMS)        // 10752 samples per window at most
#define K3_MAXP 72                        // peaks are > 150 samples apart -> at most 71 in 10752
#define K3_BEAT 320
#define K3_NONE 0xffffu

namespace {

struct K3Smem {
  double* A;            // [L]  z-scored window
  double* B;            // [L]  prefix sums / sequential smoothgrad
  double* pw_vals;      // [128]
  double* wtot;         // [16]
  double* scal;         // [16] scalars
  double* dsq;          // [K3_MAXP] squared successive RR differences
  int* ipeaks;          // [K3_MAXP]
  int* idiff;           // [K3_MAXP] RR intervals
  int* isort;           // [K3_MAXP] RR intervals, sorted
  int* iscal;           // [16]
  int* leaf_off;        // [128] NumPy-tree leaf ranges of a length-L vector
  int* leaf_size;       // [128]
  unsigned short* beg;  // [L/2 + 2]
  unsigned short* end;
  unsigned short* cand;
  unsigned char* Q;     // [L]
  unsigned char* pw_valid;  // [128]
};

__host__ __device__ inline size_t k3_smem_layout(unsigned char* base, int L, K3Smem* s) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 15) & ~(size_t)15; return base ? base + at : (unsigned char*)0; };
  K3Smem t;
  t.A = (double*)take(sizeof(double) * L);
  t.B = (double*)take(sizeof(double) * L);
  t.pw_vals = (double*)take(sizeof(double) * 128);
  t.wtot = (double*)take(sizeof(double) * 16);
  t.scal = (double*)take(sizeof(double) * 16);
  t.dsq = (double*)take(sizeof(double) * K3_MAXP);
  t.ipeaks = (int*)take(sizeof(int) * K3_MAXP);
  t.idiff = (int*)take(sizeof(int) * K3_MAXP);
  t.isort = (int*)take(sizeof(int) * K3_MAXP);
  t.iscal = (int*)take(sizeof(int) * 16);
  t.leaf_off = (int*)take(sizeof(int) * 128);
  t.leaf_size = (int*)take(sizeof(int) * 128);
  const int rmax = L / 2 + 2;
  t.beg = (unsigned short*)take(sizeof(unsigned short) * rmax);
  t.end = (unsigned short*)take(sizeof(unsigned short) * rmax);
  t.cand = (unsigned short*)take(sizeof(unsigned short) * rmax);
  t.Q = (unsigned char*)take(L);
  t.pw_valid = (unsigned char*)take(128);
  if (s) *s = t;
  return o;
}

// np.gradient(z) magnitude: interior (z[i+1]-z[i-1])/2, one-sided at the edges
__device__ __forceinline__ double absgrad(const double* __restrict__ z, int i, int L) {
  double g;
  if (i == 0) g = __dsub_rn(z[1], z[0]);
  else if (i == L - 1) g = __dsub_rn(z[L - 1], z[L - 2]);
  else g = __dmul_rn(__dsub_rn(z[i + 1], z[i - 1]), 0.5);
  return fabs(g);
}

struct SumPlain { const double* a; __device__ __forceinline__ double operator()(int i, bool) { return a[i]; } };
struct SumDev2 {
  const double* a; double m;
  __device__ __forceinline__ double operator()(int i, bool) { const double d = __dsub_rn(a[i], m); return __dmul_rn(d, d); }
};

// NumPy-order block sum with the leaf ranges tabulated (same tree as pw_block_sum)
template <class F>
__device__ __forceinline__ double block_sum_tab(F& f, int depth, const K3Smem& sm) {
  const int nslots = 1 << depth;
  const int group = threadIdx.x >> 3, g = threadIdx.x & 7, ngroups = blockDim.x >> 3;
  for (int s0 = 0; s0 < nslots; s0 += ngroups) {
    const int slot = s0 + group;
    const int off = slot < nslots ? sm.leaf_off[slot] : 0;
    const int size = slot < nslots ? sm.leaf_size[slot] : 0;
    const int c8 = size >> 3;
    double r = 0.0;
    if (c8 > 0) {
      r = f(off + g, true);
#pragma unroll 5
      for (int k = 1; k < c8; ++k) r = __dadd_rn(r, f(off + 8 * k + g, true));
    }
    double t = __shfl_xor_sync(AEW_FULL, r, 1);
    r = __dadd_rn(r, t);
    t = __shfl_xor_sync(AEW_FULL, r, 2);
    r = __dadd_rn(r, t);
    t = __shfl_xor_sync(AEW_FULL, r, 4);
    r = __dadd_rn(r, t);
    int i = c8 << 3;
    if (size < 8) { r = 0.0; i = 0; }
    for (; i < size; ++i) r = __dadd_rn(r, f(off + i, g == 0));
    if (slot < nslots && g == 0) { sm.pw_vals[slot] = r; sm.pw_valid[slot] = size > 0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) pw_warp_combine(sm.pw_vals, sm.pw_valid, depth, threadIdx.x);
  __syncthreads();
  const double r = sm.pw_vals[0];
  __syncthreads();
  return r;
}

// exclusive block scan of one double per thread (fast path; not order-sensitive)
__device__ __forceinline__ double block_excl_scan(double v, double* wtot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double o = __shfl_up_sync(AEW_FULL, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) wtot[warp] = incl;
  __syncthreads();
  double base = 0.0;
  for (int w = 0; w < warp; ++w) base += wtot[w];
  __syncthreads();
  return base + (incl - v);
}
__device__ __forceinline__ int block_excl_scan_int(int v, int* wtot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(AEW_FULL, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) wtot[warp] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += wtot[w];
  __syncthreads();
  return base + (incl - v);
}

// sum over a clamped ("nearest") window [lo, hi] from an inclusive prefix array
__device__ __forceinline__ double clamped_window(const double* __restrict__ P, int lo, int hi, int L, double first,
                                                 double last) {
  if (lo >= 1 && hi <= L - 1) return P[hi] - P[lo - 1];          // interior: no clamping
  const int nl = lo < 0 ? -lo : 0, nr = hi > L - 1 ? hi - (L - 1) : 0;
  const int lo2 = lo < 0 ? 0 : lo, hi2 = hi > L - 1 ? L - 1 : hi;
  double s = P[hi2] - (lo2 > 0 ? P[lo2 - 1] : 0.0);
  s += (double)nl * first + (double)nr * last;
  return s;
}

// ---------------------------------------------------------------------- medians
__device__ __forceinline__ void cex(double& a, double& b) {          // a <- min, b <- max (DSETP + 4 FSEL)
  const bool s = a > b;
  const double lo = s ? b : a, hi = s ? a : b;
  a = lo; b = hi;
}

// Batcher odd-even merge sort of N (power of two) values held in registers; the
// recursion is unfolded by templates so that every index is a compile-time constant
// (191 compare-exchanges for N = 32).  [LO, HI] inclusive.
template <int LO, int HI, int R>
struct OddEvenMerge {
  template <int N>
  static __device__ __forceinline__ void apply(double (&v)[N]) {
    constexpr int STEP = R * 2;
    if constexpr (STEP < HI - LO) {
      OddEvenMerge<LO, HI, STEP>::apply(v);
      OddEvenMerge<LO + R, HI, STEP>::apply(v);
#pragma unroll
      for (int i = LO + R; i < HI - R; i += STEP) cex(v[i], v[i + R]);
    } else {
      cex(v[LO], v[LO + R]);
    }
  }
};
template <int LO, int HI>
struct OddEvenSort {
  template <int N>
  static __device__ __forceinline__ void apply(double (&v)[N]) {
    if constexpr (HI - LO >= 1) {
      constexpr int MID = LO + (HI - LO) / 2;
      OddEvenSort<LO, MID>::apply(v);
      OddEvenSort<MID + 1, HI>::apply(v);
      OddEvenMerge<LO, HI, 1>::apply(v);
    }
  }
};
template <int N>
__device__ __forceinline__ void sort_network(double (&v)[N]) { OddEvenSort<0, N - 1>::apply(v); }

// value of beat b at column c (median_sample.py:30-41): short slices are LEFT-aligned
// in a NaN row; +inf stands for the NaN padding
__device__ __forceinline__ double beat_value(const double* __restrict__ z, int p, int c, int L, int pre, int post,
                                             int& valid) {
  const int s = p - pre < 0 ? 0 : p - pre;
  const int e = p + post > L ? L : p + post;
  const int idx = s + c;
  if (idx < e) { ++valid; return z[idx]; }
  return __longlong_as_double(0x7ff0000000000000ll);
}

// column c of a window with n <= 32 beats, by ONE thread: (lo + hi) / 2 with lo == hi for
// odd counts, exactly as np.ma.median evaluates it; all-NaN column -> NaN
__device__ __forceinline__ double column_median_thread(const double* __restrict__ z, const int* __restrict__ peaks,
                                                       int n, int c, int L, int pre, int post) {
  double v[32];
  int valid = 0;
#pragma unroll
  for (int b = 0; b < 32; ++b)
    v[b] = b < n ? beat_value(z, peaks[b], c, L, pre, post, valid) : __longlong_as_double(0x7ff0000000000000ll);
  sort_network<32>(v);
  const int h = valid >> 1;
  const int rhi = h < 31 ? h : 31;
  const int rlo = (valid & 1) ? h : (h > 0 ? h - 1 : 0);
  double lo = 0.0, hi = 0.0;
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    lo = (b == rlo) ? v[b] : lo;
    hi = (b == rhi) ? v[b] : hi;
  }
  if (valid == 0) return __longlong_as_double(0x7ff8000000000000ll);
  return __dmul_rn(__dadd_rn(lo, hi), 0.5);
}

// ---- warp bitonic sort of 32*K doubles, element e = k*32 + lane, ascending (n > 32 beats)
template <int K>
__device__ __forceinline__ void warp_sort(double (&v)[K], int lane) {
#pragma unroll
  for (int size = 2; size <= 32 * K; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int ks = stride >> 5;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if ((k & ks) == 0) {
            const bool asc = (((k * 32) & size) == 0);
            const double a = v[k], b = v[k | ks];
            const bool swap = asc ? (a > b) : (a < b);
            v[k] = swap ? b : a;
            v[k | ks] = swap ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const bool asc = (((k * 32 + lane) & size) == 0);
          const double o = __shfl_xor_sync(AEW_FULL, v[k], stride);
          const bool take_min = (((lane & stride) == 0) == asc);
          const bool o_lt = o < v[k];
          v[k] = (take_min == o_lt) ? o : v[k];         // equal values: either copy is the same number
        }
      }
    }
  }
}

template <int K>
__device__ __forceinline__ double warp_pick(const double (&v)[K], int rank) {
  double r = 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double c = __shfl_sync(AEW_FULL, v[k], rank & 31);
    r = ((rank >> 5) == k) ? c : r;
  }
  return r;
}

template <int K>
__device__ __forceinline__ double column_median_warp(const double* __restrict__ z, const int* __restrict__ peaks,
                                                     int n, int c, int L, int pre, int post, int lane) {
  double v[K];
  int valid = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int b = k * 32 + lane;
    v[k] = b < n ? beat_value(z, peaks[b], c, L, pre, post, valid) : __longlong_as_double(0x7ff0000000000000ll);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(AEW_FULL, valid, o);
  warp_sort<K>(v, lane);
  const int h = valid >> 1;
  const double hi = warp_pick<K>(v, h < 32 * K - 1 ? h : 32 * K - 1);
  const double lo = warp_pick<K>(v, (valid & 1) ? h : (h > 0 ? h - 1 : 0));
  if (valid == 0) return __longlong_as_double(0x7ff8000000000000ll);
  return __dmul_rn(__dadd_rn(lo, hi), 0.5);
}

// all pre+post columns of one window; every thread of the block calls it
__device__ __forceinline__ void window_median(const double* __restrict__ z, const int* __restrict__ peaks, int n,
                                              int L, int pre, int post, double* __restrict__ med,
                                              float* __restrict__ waves) {
  const int S = pre + post;
  if (n <= 32) {
    for (int c = threadIdx.x; c < S; c += blockDim.x) {
      const double m = column_median_thread(z, peaks, n, c, L, pre, post);
      if (med) med[c] = m;
      if (waves) waves[c] = (float)m;
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int c = warp; c < S; c += nwarp) {
    const double m = n <= 64 ? column_median_warp<2>(z, peaks, n, c, L, pre, post, lane)
                             : column_median_warp<4>(z, peaks, n, c, L, pre, post, lane);
    if (lane == 0) {
      if (med) med[c] = m;
      if (waves) waves[c] = (float)m;
    }
  }
}

// ---------------------------------------------------------------------- the window kernel
__global__ void __launch_bounds__(K3_NT, 1) k3_windows(K3Params p) {
  AEW_DYN_SMEM(smem_raw);
  K3Smem sm;
  k3_smem_layout(smem_raw, p.L, &sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = p.L;
  const long long w = blockIdx.x;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  if (p.win_start[w] < 0) {                             // padding slot of an over-allocated window table
    if (tid == 0) {
      p.keep[w] = 0; p.has_nan[w] = 2; p.n_peaks[w] = 0;
      for (int k = 0; k < 4; ++k) p.stats[4 * w + k] = qnan;
    }
    for (int i = tid; i < K3_MAXP; i += K3_NT) p.peaks[w * K3_MAXP + i] = -1;
    return;
  }
  const double* src = p.y + p.win_start[w];
  double* A = sm.A;
  double* B = sm.B;
  const int depth = pw_depth(L);
  if (tid < (1 << depth)) {
    long long off, size;
    pw_node(L, depth, (unsigned)tid, off, size);
    sm.leaf_off[tid] = (int)off;
    sm.leaf_size[tid] = (int)size;
  }

  // ---- load, NaN screen (remove_samples.py:31-33)
  int bad = 0;
  for (int i = tid; i < L; i += K3_NT) {
    const double v = src[i];
    A[i] = v;
    bad |= isnan(v) ? 1 : 0;
  }
  bad = __syncthreads_or(bad);
  double mean = 0.0, sd0 = 0.0;
  if (!bad && p.zscore_windows) {
    // ---- per-window z-score (data_extract_vitaldb.py:93), NumPy summation order
    SumPlain f1{A};
    mean = __ddiv_rn(block_sum_tab(f1, depth, sm), (double)L);
    SumDev2 f2{A, mean};
    sd0 = __dsqrt_rn(__ddiv_rn(block_sum_tab(f2, depth, sm), (double)L));
    int nonfinite = 0;
    const ExactDiv dv = make_exact_div(sd0);
    for (int i = tid; i < L; i += K3_NT) {
      const double z = div_exact(__dsub_rn(A[i], mean), dv);
      A[i] = z;
      nonfinite |= (isnan(z) || isinf(z)) ? 1 : 0;     // constant window: 0/0
    }
    bad = __syncthreads_or(nonfinite);
  }
  if (p.zwin) {
    double* zw = p.zwin + w * (long long)L;
    for (int i = tid; i < L; i += K3_NT) zw[i] = A[i];  // NaN windows: as z-scored by the reference or raw
  }
  if (bad) {
    if (tid == 0) {
      p.keep[w] = 0; p.has_nan[w] = 1; p.n_peaks[w] = 0;
      for (int k = 0; k < 4; ++k) p.stats[4 * w + k] = qnan;
    }
    for (int i = tid; i < K3_MAXP; i += K3_NT) p.peaks[w * K3_MAXP + i] = -1;
    return;
  }

  // ---- ecg_sd = np.std(window) (remove_samples.py:62)
  SumPlain f3{A};
  const double m2 = __ddiv_rn(block_sum_tab(f3, depth, sm), (double)L);
  SumDev2 f4{A, m2};
  const double sd = __dsqrt_rn(__ddiv_rn(block_sum_tab(f4, depth, sm), (double)L));

  // ---- QRS indicator: smoothgrad > 1.5 * avggrad
  const int ks = p.smooth_k, ka = p.avg_k;
  const int ls = ks / 2, la = ka / 2;                  // uniform_filter1d: window [i - k/2, i - k/2 + k)
  int hazard = 0;
  unsigned qbits = 0;                                  // bit k: QRS indicator of sample tid * ITEMS + k
  const int i0 = tid * K3_ITEMS;
  if (p.mode != 1) {
    double loc[K3_ITEMS], smv[K3_ITEMS];
    double run = 0.0;
    if (i0 >= 1 && i0 + K3_ITEMS < L) {                // strip away from both window edges
      const double* zp = A + i0;
      double prv = zp[-1], cur = zp[0];
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) {
        const double nxt = zp[k + 1];
        run += fabs(__dmul_rn(__dsub_rn(nxt, prv), 0.5));
        loc[k] = run;
        prv = cur; cur = nxt;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) {
        const int i = i0 + k;
        run += (i < L) ? absgrad(A, i, L) : 0.0;
        loc[k] = run;
      }
    }
    double base = block_excl_scan(run, sm.wtot);
    if (i0 + K3_ITEMS <= L) {
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) B[i0 + k] = base + loc[k];
    } else {
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) if (i0 + k < L) B[i0 + k] = base + loc[k];
    }
    if (tid == 0) { sm.scal[0] = absgrad(A, 0, L); sm.scal[1] = absgrad(A, L - 1, L); }
    __syncthreads();
    const double ag0 = sm.scal[0], agL = sm.scal[1], ptot = B[L - 1];
    const double inv_s = 1.0 / (double)ks, inv_a = 1.0 / (double)ka;
    run = 0.0;
    if (i0 - ls >= 1 && i0 + K3_ITEMS - 1 - ls + ks - 1 <= L - 1) {   // no clamping anywhere in the strip
      const double* pl = B + i0 - ls - 1;
      const double* ph = B + i0 - ls + ks - 1;
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) {
        const double s = (ph[k] - pl[k]) * inv_s;
        smv[k] = s;
        run += s;
        loc[k] = run;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) {
        const int i = i0 + k;
        double s = 0.0;
        if (i < L) s = clamped_window(B, i - ls, i - ls + ks - 1, L, ag0, agL) * inv_s;
        smv[k] = s;
        run += s;
        loc[k] = run;
        if (i == 0) sm.scal[2] = s;
        if (i == L - 1) sm.scal[3] = s;
      }
    }
    base = block_excl_scan(run, sm.wtot);               // (its barriers also end all reads of B)
    if (i0 + K3_ITEMS <= L) {
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) B[i0 + k] = base + loc[k];
    } else {
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) if (i0 + k < L) B[i0 + k] = base + loc[k];
    }
    __syncthreads();
    const double s0 = sm.scal[2], sL = sm.scal[3];
    const double tol_abs = 2e-12 * ptot * inv_s;
    if (i0 - la >= 1 && i0 + K3_ITEMS - 1 - la + ka - 1 <= L - 1) {
      const double* pl = B + i0 - la - 1;
      const double* ph = B + i0 - la + ka - 1;
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) {
        const double thr = 1.5 * ((ph[k] - pl[k]) * inv_a);
        qbits |= (unsigned)(smv[k] > thr) << k;
        hazard |= (fabs(smv[k] - thr) <= 1e-9 * (smv[k] + thr) + tol_abs) ? 1 : 0;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) {
        const int i = i0 + k;
        if (i < L) {
          const double thr = 1.5 * (clamped_window(B, i - la, i - la + ka - 1, L, s0, sL) * inv_a);
          qbits |= (unsigned)(smv[k] > thr) << k;
          hazard |= (fabs(smv[k] - thr) <= 1e-9 * (smv[k] + thr) + tol_abs) ? 1 : 0;
        }
      }
    }
    hazard = __syncthreads_or(hazard);
    if (p.mode == 2) hazard = 0;
    if (p.mode == 3) hazard = 1;                        // test hook: every window takes the fallback transition
    if (hazard && tid == 0 && p.hazard_count) atomicAdd(p.hazard_count, 1ull);
  }
  if (p.mode == 1 || hazard) {
    // SciPy's literal evaluation: one undivided running sum per box filter (appendix B.3)
    if (tid == 0) {
      const double dks = (double)ks, dka = (double)ka;
      double t = 0.0;
      for (int e = 0; e < ks; ++e) {
        int j = e - ls; j = j < 0 ? 0 : (j > L - 1 ? L - 1 : j);
        t = __dadd_rn(t, absgrad(A, j, L));
      }
      B[0] = __ddiv_rn(t, dks);
      for (int i = 1; i < L; ++i) {
        int jin = i + ks - 1 - ls, jout = i - 1 - ls;
        jin = jin > L - 1 ? L - 1 : jin;
        jout = jout < 0 ? 0 : jout;
        t = __dadd_rn(t, __dsub_rn(absgrad(A, jin, L), absgrad(A, jout, L)));
        B[i] = __ddiv_rn(t, dks);
      }
      double t2 = 0.0;
      for (int e = 0; e < ka; ++e) {
        int j = e - la; j = j < 0 ? 0 : (j > L - 1 ? L - 1 : j);
        t2 = __dadd_rn(t2, B[j]);
      }
      sm.Q[0] = B[0] > __dmul_rn(1.5, __ddiv_rn(t2, dka));
      for (int i = 1; i < L; ++i) {
        int jin = i + ka - 1 - la, jout = i - 1 - la;
        jin = jin > L - 1 ? L - 1 : jin;
        jout = jout < 0 ? 0 : jout;
        t2 = __dadd_rn(t2, __dsub_rn(B[jin], B[jout]));
        sm.Q[i] = B[i] > __dmul_rn(1.5, __ddiv_rn(t2, dka));
      }
    }
    __syncthreads();
    qbits = 0;
#pragma unroll
    for (int k = 0; k < K3_ITEMS; ++k) if (i0 + k < L) qbits |= (unsigned)(sm.Q[i0 + k] != 0) << k;
  }

  // ---- QRS runs: beg where ~q[i] & q[i+1], end where q[i] & ~q[i+1]  (ordered compaction);
  //      the flags live in registers, only each strip's first flag is exchanged
  {
    __syncthreads();                                   // everyone is done with sm.Q of a sequential pass
    sm.Q[tid] = (unsigned char)(qbits & 1u);
    __syncthreads();
    const unsigned nbr = (tid + 1 < K3_NT) ? sm.Q[tid + 1] : 0u;
    const unsigned nextq = (qbits >> 1) | (nbr << (K3_ITEMS - 1));          // flag of sample i + 1
    int nvalid = L - 1 - i0;                            // transitions exist for i < L - 1
    nvalid = nvalid < 0 ? 0 : (nvalid > K3_ITEMS ? K3_ITEMS : nvalid);
    const unsigned vmask = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
    unsigned bbits = ~qbits & nextq & vmask, ebits = qbits & ~nextq & vmask;
    const int packed = block_excl_scan_int(__popc(bbits) | (__popc(ebits) << 16), (int*)sm.wtot);
    int ob = packed & 0xffff, oe = packed >> 16;
    while (bbits) { const int k = __ffs(bbits) - 1; bbits &= bbits - 1; sm.beg[ob++] = (unsigned short)(i0 + k); }
    while (ebits) { const int k = __ffs(ebits) - 1; ebits &= ebits - 1; sm.end[oe++] = (unsigned short)(i0 + k); }
    if (tid == K3_NT - 1) { sm.iscal[0] = ob; sm.iscal[1] = oe; }
    __syncthreads();
  }
  const int nbeg = sm.iscal[0], nend = sm.iscal[1];
  // ends that precede the first beg are dropped (at most one, runs alternate)
  const int skip = (nbeg > 0 && nend > 0 && sm.end[0] <= sm.beg[0]) ? 1 : 0;
  int num = nend - skip < nbeg ? nend - skip : nbeg;
  if (nbeg == 0) num = 0;
  const unsigned short* endp = sm.end + skip;

  // ---- min_len = mean(end - beg) * 0.4 (integer sum is exact in any order)
  {
    long long part = 0;
    for (int i = tid; i < num; i += K3_NT) part += (long long)endp[i] - (long long)sm.beg[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(AEW_FULL, part, o);
    if (lane == 0) ((long long*)sm.wtot)[warp] = part;
    __syncthreads();
    if (tid == 0) {
      long long tot = 0;
      for (int k = 0; k < K3_NT / 32; ++k) tot += ((long long*)sm.wtot)[k];
      sm.scal[4] = num > 0 ? __dmul_rn(__ddiv_rn((double)tot, (double)num), 0.4) : 0.0;
    }
    __syncthreads();
  }
  const double min_len = sm.scal[4];

  // ---- per run: most prominent strict local maximum (scipy.signal.find_peaks semantics)
  for (int r = warp; r < num; r += K3_NT / 32) {
    const int b = sm.beg[r], e = endp[r];
    const int n = e - b;
    const double* d = A + b;
    double best = -1.0;
    int best_i = 0x7fffffff;
    if (!((double)n < min_len) && n <= 128) {
      // warp-cooperative evaluation (runs of at most 128 samples): lane l holds d[32 m + l]
      double v[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) v[m] = (32 * m + lane < n) ? d[32 * m + lane] : 0.0;
      unsigned peakmask[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int i = 32 * m + lane;
        bool rise = false;
        int pk = i;
        if (i >= 1 && i < n - 1 && d[i - 1] < v[m]) {       // start of a plateau (usually one sample)
          int ahead = i + 1;
          while (ahead < n - 1 && d[ahead] == v[m]) ++ahead;
          if (d[ahead] < v[m]) { rise = true; pk = (i + ahead - 1) / 2; }
        }
        // a plateau's midpoint may sit in another lane: publish the peak at its own position
        const unsigned starts = __ballot_sync(AEW_FULL, rise);
        unsigned mine_pk = 0;
        for (unsigned s = starts; s; s &= s - 1) {
          const int src = __ffs(s) - 1;
          const int q = __shfl_sync(AEW_FULL, pk, src);
          if (q == i) mine_pk = 1;
          else if ((q >> 5) != m && (q & 31) == lane) mine_pk |= 2u << (q >> 5);   // lands in another chunk
        }
        peakmask[m] = mine_pk;
      }
      // fold "lands in another chunk" marks (plateaus longer than a chunk are vanishingly rare but legal)
      unsigned pkbits[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        bool is_pk = (peakmask[m] & 1u) != 0;
#pragma unroll
        for (int o = 0; o < 4; ++o) is_pk |= (peakmask[o] & (2u << m)) != 0;
        pkbits[m] = __ballot_sync(AEW_FULL, is_pk);
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        for (unsigned s = pkbits[m]; s; s &= s - 1) {
          const int src = __ffs(s) - 1;
          const int pk = 32 * m + src;
          const double h = __shfl_sync(AEW_FULL, v[m], src);
          // nearest strictly higher sample on each side (walks of scipy's _peak_prominences stop there)
          int nl = -1, nr = n;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int i = 32 * c + lane;
            const unsigned hi = __ballot_sync(AEW_FULL, i < n && v[c] > h);
            if (hi) {
              const unsigned below = (32 * c + 31 < pk) ? hi : ((32 * c > pk) ? 0u : (hi & ((1u << (pk - 32 * c)) - 1u)));
              const unsigned above = (32 * c > pk) ? hi : ((32 * c + 31 < pk) ? 0u : (hi & ~((2u << (pk - 32 * c)) - 1u)));
              if (below) { const int q = 32 * c + 31 - __clz((int)below); nl = q > nl ? q : nl; }
              if (above) { const int q = 32 * c + __ffs((int)above) - 1; nr = q < nr ? q : nr; }
            }
          }
          double lmin = h, rmin = h;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int i = 32 * c + lane;
            if (i > nl && i <= pk && v[c] < lmin) lmin = v[c];
            if (i >= pk && i < nr && v[c] < rmin) rmin = v[c];
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double ol = __shfl_xor_sync(AEW_FULL, lmin, o), orr = __shfl_xor_sync(AEW_FULL, rmin, o);
            lmin = ol < lmin ? ol : lmin;
            rmin = orr < rmin ? orr : rmin;
          }
          const double prom = __dsub_rn(h, lmin > rmin ? lmin : rmin);
          if (prom > best) { best = prom; best_i = pk; }       // candidates come in ascending order: first max wins
        }
      }
    } else if (!((double)n < min_len)) {
      for (int i = 1 + lane; i < n - 1; i += 32) {
        if (d[i - 1] < d[i]) {
          int ahead = i + 1;
          while (ahead < n - 1 && d[ahead] == d[i]) ++ahead;
          if (d[ahead] < d[i]) {
            const int pk = (i + ahead - 1) / 2;        // plateau midpoint
            const double h = d[pk];
            double lmin = h, rmin = h;
            for (int j = pk; j >= 0 && d[j] <= h; --j) lmin = d[j] < lmin ? d[j] : lmin;
            for (int j = pk; j <= n - 1 && d[j] <= h; ++j) rmin = d[j] < rmin ? d[j] : rmin;
            const double prom = __dsub_rn(h, lmin > rmin ? lmin : rmin);
            if (prom > best || (prom == best && pk < best_i)) { best = prom; best_i = pk; }
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(AEW_FULL, best, o);
      const int oi = __shfl_xor_sync(AEW_FULL, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    if (lane == 0) sm.cand[r] = best_i == 0x7fffffff ? (unsigned short)K3_NONE : (unsigned short)(b + best_i);
  }
  __syncthreads();

  // ---- refractory chain (sequential by definition; thread 0)
  if (tid == 0) {
    int np = 0, last = 0;
    for (int r = 0; r < num; ++r) {
      const int c = sm.cand[r];
      if (c == (int)K3_NONE) continue;
      if (c - last > p.refractory) {
        if (np < K3_MAXP) sm.ipeaks[np] = c;
        ++np;
        last = c;
      }
    }
    if (np > K3_MAXP) { atomicOr(p.status, AEW_DEV_RUN_OVERFLOW); np = K3_MAXP; }
    sm.iscal[2] = np;
  }
  __syncthreads();
  const int np = sm.iscal[2];
  // remove_samples.py:53-54 keeps 0 < p < L (always true here: p > refractory)

  // ---- HR = 60 fs / median(RR) and RMSSD (remove_samples.py:85-121), cooperatively
  const int nd = np - 1;
  if (tid < nd) {
    const int d0 = sm.ipeaks[tid + 1] - sm.ipeaks[tid];
    sm.idiff[tid] = d0;
    if (tid + 1 < nd) {
      const double cur = __ddiv_rn((double)(sm.ipeaks[tid + 2] - sm.ipeaks[tid + 1]), p.rr_div);
      const double prv = __ddiv_rn((double)d0, p.rr_div);
      const double dd = __dsub_rn(cur, prv);
      sm.dsq[tid] = __dmul_rn(dd, dd);
    }
  }
  __syncthreads();
  if (tid < nd) {                                        // rank sort of <= 71 integers
    const int mine = sm.idiff[tid];
    int rank = 0;
    for (int j = 0; j < nd; ++j) {
      const int o = sm.idiff[j];
      rank += (o < mine) || (o == mine && j < tid);
    }
    sm.isort[rank] = mine;
  }
  double ssq = 0.0;
  if (warp == 15) {                                      // NumPy leaf order over np-2 <= 70 values
    SumPlain fq{sm.dsq};
    ssq = pw_leaf_sum8(fq, 0, np >= 3 ? np - 2 : 0, lane & 7);
    if (lane == 0) sm.scal[5] = ssq;
  }
  __syncthreads();
  if (tid == 0) {
    double hr = qnan, rms = qnan;
    if (np >= 2) {
      const double med = (nd & 1) ? (double)sm.isort[nd / 2]
                                  : __dmul_rn(__dadd_rn((double)sm.isort[nd / 2 - 1], (double)sm.isort[nd / 2]), 0.5);
      hr = __ddiv_rn(p.hr_num, med);
    }
    if (np >= 3) rms = __dsqrt_rn(__ddiv_rn(sm.scal[5], (double)(np - 2)));
    const double sdn = __ddiv_rn(sd, (double)np);
    const bool remove = (hr < p.hr_lo) || (hr > p.hr_hi) || isnan(hr) || (rms > p.pvc_threshold) ||
                        (sdn > p.sd_threshold);
    sm.iscal[3] = remove ? 0 : 1;
    p.keep[w] = remove ? 0 : 1;
    p.has_nan[w] = 0;
    p.n_peaks[w] = np;
    p.stats[4 * w + 0] = hr; p.stats[4 * w + 1] = rms; p.stats[4 * w + 2] = sd; p.stats[4 * w + 3] = sdn;
  }
  for (int i = tid; i < K3_MAXP; i += K3_NT) p.peaks[w * K3_MAXP + i] = i < np ? sm.ipeaks[i] : -1;
  __syncthreads();

  // ---- median beat of kept windows (median_sample.py:6-54) -> np_waves row
  if (sm.iscal[3] && (p.median || p.waves))
    window_median(A, sm.ipeaks, np, L, p.pre, p.post, p.median ? p.median + w * K3_BEAT : nullptr,
                  p.waves ? p.waves + w * K3_BEAT : nullptr);
}

// Stand-alone median beats for given peaks (the compute_median_sample_ecg shim):
// one block per window, samples read from global memory.
__global__ void __launch_bounds__(K3_NT, 1) k4_median(const double* __restrict__ windows, int L,
                                                      const int* __restrict__ peaks,
                                                      const long long* __restrict__ peak_off, int pre, int post,
                                                      double* __restrict__ median, double* __restrict__ aligned) {
  const long long w = blockIdx.x;
  const double* z = windows + w * (long long)L;
  const int* pk = peaks + peak_off[w];
  const int n = (int)(peak_off[w + 1] - peak_off[w]);
  const int S = pre + post;
  if (aligned) {
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    double* al = aligned + peak_off[w] * S;
    for (int i = threadIdx.x; i < n * S; i += blockDim.x) {
      const int b = i / S, c = i - b * S;
      const int p = pk[b];
      const int s = p - pre < 0 ? 0 : p - pre, e = p + post > L ? L : p + post;
      al[i] = (s + c < e) ? z[s + c] : qnan;
    }
  }
  window_median(z, pk, n, L, pre, post, median + w * S, nullptr);
}

}  // namespace

extern "C" int aew_windows_max_peaks(void) { return K3_MAXP; }
extern "C" int aew_windows_max_len(void) { return K3_MAXL; }

extern "C" int aew_windows_beats_f64(const K3Params* hp, void* stream) {
  if (!hp) return AEW_ERR_ARG;
  K3Params p = *hp;
  if (!p.y || !p.win_start || !p.keep || !p.has_nan || !p.n_peaks || !p.peaks || !p.stats || !p.status)
    return AEW_ERR_ARG;
  if (p.n_win == 0) return AEW_OK;
  if (p.L < 16 || p.L > K3_MAXL || p.n_win < 0 || p.n_win > 0x7fffffffll) return AEW_ERR_UNSUPPORTED;
  if (pw_depth(p.L) > 7 || p.pre + p.post != K3_BEAT || p.smooth_k < 1 || p.avg_k < 1 || p.smooth_k > p.L ||
      p.avg_k > p.L || p.refractory < 1 || (p.L / (p.refractory + 1) + 1) > K3_MAXP)
    return AEW_ERR_UNSUPPORTED;
  const size_t smem = k3_smem_layout(nullptr, p.L, nullptr);
  if (smem > 227 * 1024) return AEW_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) { AEW_SET_MAX_SMEM(k3_windows, 227 * 1024); attr_done = true; }
  AEW_LAUNCH(k3_windows, (unsigned)p.n_win, K3_NT, smem, (cudaStream_t)stream, p);
  return (int)cudaGetLastError();
}

// windows: [n_win][L] contiguous, peaks: CSR (peak_off[n_win+1]); at most 128 peaks per window.
extern "C" int aew_median_beats_f64(const double* windows, long long n_win, int L, const int* peaks,
                                    const long long* d_peak_off, int max_peaks_per_window, int pre, int post,
                                    double* median, double* aligned, void* stream) {
  if (!windows || !peaks || !d_peak_off || !median || n_win < 0 || L < 1) return AEW_ERR_ARG;
  if (max_peaks_per_window > 128 || pre + post != K3_BEAT || n_win > 0x7fffffffll) return AEW_ERR_UNSUPPORTED;
  if (n_win == 0) return AEW_OK;
  AEW_LAUNCH(k4_median, (unsigned)n_win, K3_NT, 0, (cudaStream_t)stream, windows, L, peaks, d_peak_off, pre, post,
             median, aligned);
  return (int)cudaGetLastError();
}
