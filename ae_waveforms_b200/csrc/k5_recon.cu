// k5_recon.cu -- Gaussian fiducial-point reconstruction and the sigma-map search
// objective, fused (K5).
//
// Replaces functions_traditional/signal_reconstruction.py:5-81 (gaussian_line,
// extract_coordinates, reconstruct_gaussian) and the objective of
// analysis_performance_traditional_vs_ae.py:63-90 evaluated for MANY sigma maps at
// once: for every beat and every candidate map the beat is rebuilt as a chain of
// half-Gaussian segments between consecutive fiducial points on the integer-ms
// grid, resampled to the 320 beat samples with np.interp semantics, cast to
// float32, and the squared float32 error against the float32 target is
// accumulated (float64) per candidate.
//
// Mapping: persistent 320-thread blocks walk the beats; thread k owns output
// sample k.  Everything that does not depend on sigma -- the 13-point coordinate
// extraction (NaN drop, stable sort, duplicate bump: extract_coordinates), the
// sigma_map slot walk with its NaN quirk, which two grid
// points bracket sample k and which segments they belong to -- is computed once
// per beat; per candidate the block evaluates 2 exps per segment (the affine
// rescale of gaussian_line) and 2 exps per thread.  Candidates are processed four
// at a time to amortise the two barriers.  Per-candidate sums go to a block-private
// row of the workspace (no atomics) and a second kernel adds the rows in a fixed
// order, so the result is deterministic.
//
// Bound: FP64/exp issue, not HBM -- ~664 double-precision exp per beat.candidate
// against 1.46 KB read per beat.  No tensor cores: nothing
// here is a contraction.
#include "aew_rt.h"

#define K5_NT 320
#define K5_NPT 13            // 11 fiducials + 2 baseline end points
#define K5_NSEG 12
#define K5_CT 4              // candidates per inner iteration

namespace {

struct BeatPlan {
  double x0[K5_NSEG], y0[K5_NSEG], x1[K5_NSEG], y1[K5_NSEG];
  int slot[K5_NSEG];       // sigma_map index used by the segment
  int forced[K5_NSEG];     // 1: divisor forced to 1 (next slot was NaN)
  int emit[K5_NSEG];       // segment contributes samples
  int start[K5_NSEG];      // first index in the concatenated signal
  int gx0[K5_NSEG];        // int(x0): first grid value
  int npts[K5_NSEG];
  int nseg;                // compacted pairs
  int total;               // len(ecg_syn2)
  int invalid;             // the reference would raise (index error / empty signal)
};

struct SegCand {           // per (segment, candidate): the affine rescale of gaussian_line
  double inv2s2;           // 1 / (2 s^2)
  double amp;              // a = max(y1, y2)
  double mu;
  double c0, num, den, off;
};

__device__ void build_plan(const double* fx, const double* fy, double baseline, const double* xt, int nt,
                           BeatPlan& P) {
  double xs[K5_NPT], ys[K5_NPT];
  bool missing[K5_NPT];
  double cx[K5_NPT], cy[K5_NPT];
  xs[0] = xt[0]; ys[0] = baseline;
  for (int i = 0; i < 11; ++i) { xs[i + 1] = fx[i]; ys[i + 1] = fy[i]; }
  xs[12] = xt[nt - 1]; ys[12] = baseline;
  int m = 0;
  for (int i = 0; i < K5_NPT; ++i) {
    missing[i] = isnan(xs[i]) || isnan(ys[i]);
    if (!missing[i]) { cx[m] = xs[i]; cy[m] = ys[i]; ++m; }
  }
  // np.argsort on <= 16 elements is an insertion sort: stable
  for (int i = 1; i < m; ++i) {
    const double vx = cx[i], vy = cy[i];
    int j = i - 1;
    while (j >= 0 && cx[j] > vx) { cx[j + 1] = cx[j]; cy[j + 1] = cy[j]; --j; }
    cx[j + 1] = vx; cy[j + 1] = vy;
  }
  // duplicate x: every occurrence after the first gets += step, until unique (no re-sort)
  const double step = xt[1] - xt[0];
  for (int guard = 0; guard < 64; ++guard) {
    double dup[K5_NPT];
    int nd = 0;
    for (int i = 0; i < m; ++i) {                      // duplicated values of this round, ascending
      int count = 0;
      bool seen = false;
      for (int j = 0; j < m; ++j) count += (cx[j] == cx[i]);
      for (int j = 0; j < nd; ++j) seen |= (dup[j] == cx[i]);
      if (count > 1 && !seen) {
        int at = nd++;
        while (at > 0 && dup[at - 1] > cx[i]) { dup[at] = dup[at - 1]; --at; }
        dup[at] = cx[i];
      }
    }
    if (nd == 0) break;
    for (int d = 0; d < nd; ++d) {
      bool first = true;
      for (int i = 0; i < m; ++i) {
        if (cx[i] == dup[d]) {
          if (!first) cx[i] += step;
          first = false;
        }
      }
    }
  }
  // segment chain with the sigma_map slot walk (signal_reconstruction.py:22-36)
  P.invalid = 0;
  P.nseg = m > 0 ? m - 1 : 0;
  int slot = 0, total = 0;
  const double lo = xt[0], hi = xt[nt - 1];
  for (int i = 0; i + 1 < m; ++i) {
    P.x0[i] = cx[i]; P.y0[i] = cy[i]; P.x1[i] = cx[i + 1]; P.y1[i] = cy[i + 1];
    if (slot > 11) { P.invalid = 1; slot = 11; }      // sigma_map[slot] would raise IndexError
    P.slot[i] = slot;
    P.forced[i] = 0;
    while (true) {
      if (slot + 1 > 12) { P.invalid = 1; break; }    // nan_mask[slot + 1] would raise IndexError
      if (!missing[slot + 1]) break;
      P.forced[i] = 1;
      ++slot;
    }
    const bool emit = (cx[i + 1] - cx[i] > 1.0) && (cx[i + 1] > lo) && (cx[i] < hi);
    const int g0 = (int)cx[i], g1 = (int)cx[i + 1];   // int(): truncation toward zero
    const int np = emit && g1 > g0 ? g1 - g0 : 0;
    if (emit && np <= 0) P.invalid = 1;           // gaussian_line on an empty grid raises in the reference
    P.emit[i] = emit && np > 0;
    P.gx0[i] = g0;
    P.npts[i] = np;
    P.start[i] = total;
    total += np;
    ++slot;
  }
  P.total = total;
  if (total == 0) P.invalid = 1;                       // np.interp on an empty signal raises
}

__global__ void __launch_bounds__(K5_NT) k5_recon(const double* __restrict__ fx, const double* __restrict__ fy,
                                                  const double* __restrict__ baseline, const float* __restrict__ target,
                                                  long long n_beats, const double* __restrict__ xt, int nt,
                                                  const double* __restrict__ cand, long long n_cand,
                                                  double* __restrict__ partial, double* __restrict__ pred_out) {
  __shared__ BeatPlan P;
  __shared__ SegCand SC[K5_CT][K5_NSEG];
  __shared__ double s_red[K5_CT][K5_NT / 32];
  __shared__ double s_sig[K5_CT][12];
  const int k = threadIdx.x, lane = k & 31, warp = k >> 5;
  double* my_row = partial + (long long)blockIdx.x * n_cand;
  for (long long c = k; c < n_cand; c += K5_NT) my_row[c] = 0.0;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  const double x = k < nt ? xt[k] : 0.0;

  for (long long beat = blockIdx.x; beat < n_beats; beat += gridDim.x) {
    __syncthreads();
    if (k == 0) build_plan(fx + 11 * beat, fy + 11 * beat, baseline[beat], xt, nt, P);
    __syncthreads();
    const float tgt = (k < nt && target) ? target[beat * nt + k] : 0.f;
    // ---- sigma-independent placement of output sample k (np.interp over np.linspace)
    int ja = -1, jb = -1, sa = 0, sb = 0;             // grid indices / segments of the two bracketing points
    double dxa = 0.0, dxb = 0.0;                      // grid x minus mu of the segment
    double t1 = 0.0, den = 1.0;
    const int total = P.total;
    if (!P.invalid && k < nt) {
      const double start = xt[0], stop = xt[nt - 1];
      if (total == 1) { ja = 0; }
      else {
        const double stepv = (stop - start) / (double)(total - 1);
        auto xp = [&](int j) { return j == total - 1 ? stop : (double)j * stepv + start; };
        int j = (int)floor((x - start) / stepv);
        j = j < 0 ? 0 : (j > total - 1 ? total - 1 : j);
        while (j > 0 && xp(j) > x) --j;
        while (j + 1 < total && xp(j + 1) <= x) ++j;
        ja = j;
        if (j < total - 1 && xp(j) != x) { jb = j + 1; t1 = x - xp(j); den = xp(j + 1) - xp(j); }
      }
      for (int s = 0; s < P.nseg; ++s) {
        if (!P.emit[s]) continue;
        if (ja >= P.start[s] && ja < P.start[s] + P.npts[s]) sa = s;
        if (jb >= P.start[s] && jb < P.start[s] + P.npts[s]) sb = s;
      }
      const double mua = P.y1[sa] > P.y0[sa] ? P.x1[sa] : P.x0[sa];
      const double mub = P.y1[sb] > P.y0[sb] ? P.x1[sb] : P.x0[sb];
      dxa = (double)(P.gx0[sa] + (ja - P.start[sa])) - mua;
      dxb = (double)(P.gx0[sb] + (jb - P.start[sb])) - mub;
    }

    for (long long c0 = 0; c0 < n_cand; c0 += K5_CT) {
      const int nc = (int)(n_cand - c0 < K5_CT ? n_cand - c0 : K5_CT);
      if (k < K5_CT * 12) {
        const int ci = k / 12, si = k - 12 * ci;
        if (ci < nc) s_sig[ci][si] = cand[(c0 + ci) * 12 + si];
      }
      __syncthreads();
      // ---- per (candidate, segment): end values of the raw Gaussian and the rescale
      if (k < K5_CT * K5_NSEG) {
        const int ci = k / K5_NSEG, s = k - K5_NSEG * ci;
        if (ci < nc && s < P.nseg && P.emit[s]) {
          const double xa = P.x0[s], xb = P.x1[s], ya = P.y0[s], yb = P.y1[s];
          const double width = fabs(xb - xa);
          const double sg = P.forced[s] ? width / 1.0 : width / s_sig[ci][P.slot[s]];
          const double amp = ya > yb ? ya : yb;          // max(y1, y2)
          const bool rising = yb > ya;
          const double mu = rising ? xb : xa;
          const double two_s2 = 2.0 * (sg * sg);
          const double gfirst_x = (double)P.gx0[s] - mu, glast_x = (double)(P.gx0[s] + P.npts[s] - 1) - mu;
          const double gf = amp * exp(-(gfirst_x * gfirst_x) / two_s2);
          const double gl = amp * exp(-(glast_x * glast_x) / two_s2);
          SegCand q;
          q.inv2s2 = two_s2; q.amp = amp; q.mu = mu;
          if (rising) { q.c0 = gf; q.num = yb - ya; q.den = gl - gf; q.off = ya; }
          else { q.c0 = gl; q.num = ya - yb; q.den = gf - gl; q.off = yb; }
          SC[ci][s] = q;
        }
      }
      __syncthreads();
      // ---- my sample for each candidate of the tile
      double acc[K5_CT];
#pragma unroll
      for (int ci = 0; ci < K5_CT; ++ci) {
        acc[ci] = 0.0;
        if (ci < nc && k < nt) {
          double pred = qnan;
          if (!P.invalid) {
            const SegCand& qa = SC[ci][sa];
            const double ga = qa.amp * exp(-(dxa * dxa) / qa.inv2s2);
            const double fa = (ga - qa.c0) * qa.num / qa.den + qa.off;
            pred = fa;
            if (jb >= 0) {
              const SegCand& qb = SC[ci][sb];
              const double gb = qb.amp * exp(-(dxb * dxb) / qb.inv2s2);
              const double fb = (gb - qb.c0) * qb.num / qb.den + qb.off;
              const double slope = (fb - fa) / den;
              pred = slope * t1 + fa;
              if (isnan(pred)) {                         // np.interp's NaN rescue
                pred = slope * (t1 - den) + fb;
                if (isnan(pred) && fa == fb) pred = fa;
              }
            }
          }
          if (pred_out && c0 + ci == 0) pred_out[beat * nt + k] = pred;
          const float d = __fsub_rn((float)pred, tgt);
          acc[ci] = (double)__fmul_rn(d, d);
        }
      }
#pragma unroll
      for (int ci = 0; ci < K5_CT; ++ci) {
        double v = acc[ci];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(AEW_FULL, v, o);
        if (lane == 0) s_red[ci][warp] = v;
      }
      __syncthreads();
      if (k < nc) {
        double v = 0.0;
        for (int w2 = 0; w2 < K5_NT / 32; ++w2) v += s_red[k][w2];
        my_row[c0 + k] += v;
      }
    }
  }
}

__global__ void k5_reduce(const double* __restrict__ partial, int n_rows, long long n_cand, double* __restrict__ sse) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  double s = 0.0;
  for (int r = 0; r < n_rows; ++r) s += partial[(long long)r * n_cand + c];
  sse[c] = s;
}

// ------------------------------------------------------------------------------------------------
// Separable formulation for DISCRETE sigma grids (the reference's own search space is Integer(1,5)^12,
// analysis_performance_traditional_vs_ae.py:58).  The resampling grid of a beat does not depend on sigma,
// so output sample k depends on the sigma of at most two segments (the segments of its two bracketing grid
// points).  Per beat the kernel evaluates every sample for every grid value of the one or two slots it
// depends on and accumulates, over all beats,
//     T0              samples that depend on no slot (forced divisor 1)
//     T1[slot][v]     samples whose two points lie in one segment reading sigma_map[slot]
//     T2[a][b][v][w]  samples bracketed by two segments reading slots a and b
// after which any candidate costs 1 + 12 + 144 table look-ups:  SSE(c) = T0 + sum_s T1[s][c_s] + sum_ab T2[a][b][c_a][c_b].
// Same per-sample arithmetic as the brute-force kernel (identical float32 squared errors), ~700x less work
// for 4,096 candidates.  Up to K5_NV grid values.
#define K5_NV 5
#define K5_T1 (12 * K5_NV)
#define K5_T2 (144 * K5_NV * K5_NV)
#define K5_TABLE (1 + K5_T1 + K5_T2)

struct SamplePos {          // sigma-independent placement of one output sample (np.interp over np.linspace)
  int ja, jb, sa, sb;
  double dxa, dxb, t1, den;
};

__device__ __forceinline__ SamplePos place_sample(const BeatPlan& P, const double* __restrict__ xt, int nt, int k) {
  SamplePos q;
  q.ja = -1; q.jb = -1; q.sa = 0; q.sb = 0; q.dxa = 0.0; q.dxb = 0.0; q.t1 = 0.0; q.den = 1.0;
  const int total = P.total;
  if (P.invalid || k >= nt) return q;
  const double x = xt[k], start = xt[0], stop = xt[nt - 1];
  if (total == 1) { q.ja = 0; }
  else {
    const double stepv = (stop - start) / (double)(total - 1);
    auto xp = [&](int j) { return j == total - 1 ? stop : (double)j * stepv + start; };
    int j = (int)floor((x - start) / stepv);
    j = j < 0 ? 0 : (j > total - 1 ? total - 1 : j);
    while (j > 0 && xp(j) > x) --j;
    while (j + 1 < total && xp(j + 1) <= x) ++j;
    q.ja = j;
    if (j < total - 1 && xp(j) != x) { q.jb = j + 1; q.t1 = x - xp(j); q.den = xp(j + 1) - xp(j); }
  }
  for (int s = 0; s < P.nseg; ++s) {
    if (!P.emit[s]) continue;
    if (q.ja >= P.start[s] && q.ja < P.start[s] + P.npts[s]) q.sa = s;
    if (q.jb >= P.start[s] && q.jb < P.start[s] + P.npts[s]) q.sb = s;
  }
  const double mua = P.y1[q.sa] > P.y0[q.sa] ? P.x1[q.sa] : P.x0[q.sa];
  const double mub = P.y1[q.sb] > P.y0[q.sb] ? P.x1[q.sb] : P.x0[q.sb];
  q.dxa = (double)(P.gx0[q.sa] + (q.ja - P.start[q.sa])) - mua;
  q.dxb = (double)(P.gx0[q.sb] + (q.jb - P.start[q.sb])) - mub;
  return q;
}

// affine rescale of gaussian_line for segment s of the plan with divisor `div` (sigma = |dx| / div)
__device__ __forceinline__ SegCand seg_constants(const BeatPlan& P, int s, double div) {
  const double xa = P.x0[s], xb = P.x1[s], ya = P.y0[s], yb = P.y1[s];
  const double width = fabs(xb - xa);
  const double sg = width / div;
  const double amp = ya > yb ? ya : yb;          // max(y1, y2)
  const bool rising = yb > ya;
  const double mu = rising ? xb : xa;
  const double two_s2 = 2.0 * (sg * sg);
  const double gfirst_x = (double)P.gx0[s] - mu, glast_x = (double)(P.gx0[s] + P.npts[s] - 1) - mu;
  const double gf = amp * exp(-(gfirst_x * gfirst_x) / two_s2);
  const double gl = amp * exp(-(glast_x * glast_x) / two_s2);
  SegCand q;
  q.inv2s2 = two_s2; q.amp = amp; q.mu = mu;
  if (rising) { q.c0 = gf; q.num = yb - ya; q.den = gl - gf; q.off = ya; }
  else { q.c0 = gl; q.num = ya - yb; q.den = gf - gl; q.off = yb; }
  return q;
}

__device__ __forceinline__ double eval_point(const SegCand& q, double dx) {
  const double g = q.amp * exp(-(dx * dx) / q.inv2s2);
  return (g - q.c0) * q.num / q.den + q.off;
}

// np.interp between two evaluated grid points (or the single point), incl. its NaN rescue
__device__ __forceinline__ double interp_pred(const SamplePos& q, double fa, double fb) {
  if (q.jb < 0) return fa;
  const double slope = (fb - fa) / q.den;
  double pred = slope * q.t1 + fa;
  if (isnan(pred)) {
    pred = slope * (q.t1 - q.den) + fb;
    if (isnan(pred) && fa == fb) pred = fa;
  }
  return pred;
}

__device__ __forceinline__ double sq_err_f32(double pred, float tgt) {
  const float d = __fsub_rn((float)pred, tgt);
  return (double)__fmul_rn(d, d);
}

__global__ void __launch_bounds__(K5_NT) k5_tables(const double* __restrict__ fx, const double* __restrict__ fy,
                                                   const double* __restrict__ baseline, const float* __restrict__ target,
                                                   long long n_beats, const double* __restrict__ xt, int nt,
                                                   const double* __restrict__ vals, int nv, double* __restrict__ partial) {
  __shared__ BeatPlan P;
  __shared__ SegCand SC[K5_NSEG][K5_NV];               // [segment][grid value] (forced segments: entry 0 only)
  __shared__ double s_err[K5_NT][K5_NV];               // per sample: squared error for each value of ITS slot
  __shared__ int s_slot[K5_NT];                        // slot the sample depends on: -1 none, -2 two slots, -3 unused
  __shared__ double T2[K5_T2];
  const int k = threadIdx.x;
  for (int i = k; i < K5_T2; i += K5_NT) T2[i] = 0.0;
  double t1acc = 0.0;                                  // thread (slot, v) = k < 12 * nv accumulates T1[slot][v]; thread 12*nv: T0
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  const int t2_first = 96;                             // threads [96, 96 + nv * nv) own the (v, w) cells of T2

  for (long long beat = blockIdx.x; beat < n_beats; beat += gridDim.x) {
    __syncthreads();
    if (k == 0) build_plan(fx + 11 * beat, fy + 11 * beat, baseline[beat], xt, nt, P);
    __syncthreads();
    if (k < K5_NSEG * K5_NV) {                         // rescale constants per (segment, grid value)
      const int s = k / K5_NV, v = k - K5_NV * s;
      if (s < P.nseg && P.emit[s] && v < nv && (!P.forced[s] || v == 0))
        SC[s][v] = seg_constants(P, s, P.forced[s] ? 1.0 : vals[v]);
    }
    const SamplePos q = place_sample(P, xt, nt, k);
    const float tgt = (k < nt) ? target[beat * nt + k] : 0.f;
    __syncthreads();
    int slot = -3;
    if (k < nt) {
      if (P.invalid) {                                 // the reference raises for this beat: every candidate gets NaN
        slot = -1;
        s_err[k][0] = qnan;
      } else {
        const bool has_b = q.jb >= 0;
        const bool free_a = !P.forced[q.sa];
        const bool free_b = has_b && !P.forced[q.sb];
        if (free_a && free_b && q.sa != q.sb) {
          slot = -2;                                   // two candidate entries: T2 threads below
        } else if (free_a || free_b) {
          slot = free_a ? P.slot[q.sa] : P.slot[q.sb];
          for (int v = 0; v < nv; ++v) {
            const double fa = eval_point(SC[q.sa][free_a ? v : 0], q.dxa);
            const double fb = has_b ? eval_point(SC[q.sb][free_b ? v : 0], q.dxb) : fa;
            s_err[k][v] = sq_err_f32(interp_pred(q, fa, fb), tgt);
          }
        } else {
          slot = -1;
          const double fa = eval_point(SC[q.sa][0], q.dxa);
          const double fb = has_b ? eval_point(SC[q.sb][0], q.dxb) : fa;
          s_err[k][0] = sq_err_f32(interp_pred(q, fa, fb), tgt);
        }
      }
    }
    s_slot[k] = slot;
    __syncthreads();
    // every accumulator walks the samples in index order, so the sums do not depend on thread timing
    if (k < 12 * nv) {
      const int sl = k / nv, v = k - nv * sl;
      for (int i = 0; i < nt; ++i) if (s_slot[i] == sl) t1acc += s_err[i][v];
    } else if (k == 12 * nv) {
      for (int i = 0; i < nt; ++i) if (s_slot[i] == -1) t1acc += s_err[i][0];
    } else if (k >= t2_first && k < t2_first + nv * nv) {
      const int v = (k - t2_first) / nv, w = (k - t2_first) - nv * v;
      for (int i = 0; i < nt; ++i) {
        if (s_slot[i] != -2) continue;
        const SamplePos qi = place_sample(P, xt, nt, i);
        const double fa = eval_point(SC[qi.sa][v], qi.dxa);
        const double fb = eval_point(SC[qi.sb][w], qi.dxb);
        const double e = sq_err_f32(interp_pred(qi, fa, fb), target[beat * nt + i]);
        T2[((P.slot[qi.sa] * 12 + P.slot[qi.sb]) * K5_NV + v) * K5_NV + w] += e;
      }
    }
  }
  __syncthreads();
  double* out = partial + (long long)blockIdx.x * K5_TABLE;   // zero-filled by the host: unused columns stay 0
  if (k < 12 * nv) out[1 + (k / nv) * K5_NV + (k % nv)] = t1acc;
  else if (k == 12 * nv) out[0] = t1acc;
  for (int i = k; i < K5_T2; i += K5_NT) out[1 + K5_T1 + i] = T2[i];
}

// tables[i] = sum over blocks, fixed order
__global__ void k5_tables_reduce(const double* __restrict__ partial, int n_rows, double* __restrict__ tables) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K5_TABLE) return;
  double s = 0.0;
  for (int r = 0; r < n_rows; ++r) s += partial[(long long)r * K5_TABLE + i];
  tables[i] = s;
}

// SSE of candidate c from the tables; idx[c][s] = index of sigma_map[s] in the value grid
__global__ void k5_score(const double* __restrict__ tables, const unsigned char* __restrict__ idx, long long n_cand,
                         double* __restrict__ sse) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const unsigned char* ci = idx + 12 * c;
  double s = tables[0];
  for (int a = 0; a < 12; ++a) s += tables[1 + a * K5_NV + ci[a]];
  const double* t2 = tables + 1 + K5_T1;
  for (int a = 0; a < 12; ++a)
    for (int b = 0; b < 12; ++b) s += t2[((a * 12 + b) * K5_NV + ci[a]) * K5_NV + ci[b]];
  sse[c] = s;
}

// gaussian_line (signal_reconstruction.py:5-14) for one segment: out[i], i < int(x2)-int(x1)
__global__ void k5_gaussian_line(double x1, double y1, double x2, double y2, double s, double* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = y1 > y2 ? y1 : y2;
  const bool rising = y2 > y1;
  const double mu = rising ? x2 : x1;
  const int g0 = (int)x1;
  const double two_s2 = 2.0 * (s * s);
  auto g = [&](int j) { const double d = (double)(g0 + j) - mu; return a * exp(-(d * d) / two_s2); };
  const double gi = g(i), gf = g(0), gl = g(n - 1);
  out[i] = rising ? (gi - gf) * (y2 - y1) / (gl - gf) + y1 : (gi - gl) * (y1 - y2) / (gf - gl) + y2;
}

// extract_coordinates (signal_reconstruction.py:45-81) for one beat: coords[0..m) x, coords[13..13+m) y,
// meta[0] = m, meta[1..14) = nan mask
__global__ void k5_extract(const double* __restrict__ fx, const double* __restrict__ fy, double baseline,
                           const double* __restrict__ xt, int nt, double* __restrict__ coords, int* __restrict__ meta) {
  __shared__ BeatPlan P;
  if (threadIdx.x != 0) return;
  build_plan(fx, fy, baseline, xt, nt, P);
  const int m = P.nseg + 1;
  int cnt = 0;
  for (int i = 0; i < K5_NPT; ++i) {
    const double xv = i == 0 ? xt[0] : (i == 12 ? xt[nt - 1] : fx[i - 1]);
    const double yv = (i == 0 || i == 12) ? baseline : fy[i - 1];
    const int miss = isnan(xv) || isnan(yv);
    meta[1 + i] = miss;
    cnt += !miss;
  }
  meta[0] = cnt;
  if (cnt >= 2) {
    for (int i = 0; i < m - 1; ++i) { coords[i] = P.x0[i]; coords[13 + i] = P.y0[i]; }
    coords[m - 1] = P.x1[m - 2];
    coords[13 + m - 1] = P.y1[m - 2];
  } else if (cnt == 1) {
    for (int i = 0; i < K5_NPT; ++i)
      if (!meta[1 + i]) {
        coords[0] = i == 0 ? xt[0] : (i == 12 ? xt[nt - 1] : fx[i - 1]);
        coords[13] = (i == 0 || i == 12) ? baseline : fy[i - 1];
      }
  }
}

int k5_grid(long long n_beats) {
#ifdef AEW_EMUL
  const long long cap = 2;
#else
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 4;
#endif
  return (int)(n_beats < cap ? (n_beats > 0 ? n_beats : 1) : cap);
}

}  // namespace

extern "C" size_t aew_recon_workspace_bytes(long long n_beats, long long n_cand) {
  return (size_t)k5_grid(n_beats) * (size_t)n_cand * sizeof(double) + 256;
}

// fx, fy: [n_beats][11] float64 (NaN = missing fiducial) in the order of
// analysis_performance_traditional_vs_ae.py:40-44; baseline [n_beats]; target
// [n_beats][n_t] float32 (may be NULL when only d_pred is wanted); xt [n_t] beat time
// axis in ms; cand [n_cand][12] sigma maps; sse [n_cand] out: sum over beats and
// samples of (float32(pred) - target)^2.  d_pred (optional) receives the float64
// reconstructions [n_beats][n_t] of candidate 0.
extern "C" int aew_recon_sse_f64(const double* d_fx, const double* d_fy, const double* d_baseline,
                                 const float* d_target, long long n_beats, const double* d_xt, int n_t,
                                 const double* d_cand, long long n_cand, double* d_sse, double* d_pred,
                                 void* d_ws, size_t ws_bytes, void* stream) {
  if (!d_fx || !d_fy || !d_baseline || !d_xt || !d_cand || !d_sse || !d_ws || n_beats < 0 || n_cand < 1)
    return AEW_ERR_ARG;
  if (n_t < 2 || n_t > K5_NT) return AEW_ERR_UNSUPPORTED;
  if (ws_bytes < aew_recon_workspace_bytes(n_beats, n_cand)) return AEW_ERR_WORKSPACE;
  const int grid = k5_grid(n_beats);
  double* partial = (double*)(((size_t)d_ws + 127) & ~(size_t)127);
  cudaStream_t st = (cudaStream_t)stream;
  AEW_LAUNCH(k5_recon, grid, K5_NT, 0, st, d_fx, d_fy, d_baseline, d_target, n_beats, d_xt, n_t, d_cand, n_cand,
             partial, d_pred);
  AEW_LAUNCH(k5_reduce, (unsigned)((n_cand + 127) / 128), 128, 0, st, (const double*)partial, grid, n_cand, d_sse);
  return (int)cudaGetLastError();
}

// gaussian_line(x1, y1, x2, y2, s): n = int(x2) - int(x1) values into d_out
extern "C" int aew_gaussian_line_f64(double x1, double y1, double x2, double y2, double s, double* d_out, int n,
                                     void* stream) {
  if (!d_out || n < 0) return AEW_ERR_ARG;
  if (n == 0) return AEW_OK;
  AEW_LAUNCH(k5_gaussian_line, (unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream, x1, y1, x2, y2, s, d_out, n);
  return (int)cudaGetLastError();
}

// extract_coordinates for one beat: d_fx, d_fy [11]; d_coords [26] out (x then y), d_meta [14] out
extern "C" int aew_extract_coordinates_f64(const double* d_fx, const double* d_fy, double baseline,
                                           const double* d_xt, int n_t, double* d_coords, int* d_meta, void* stream) {
  if (!d_fx || !d_fy || !d_xt || !d_coords || !d_meta || n_t < 2) return AEW_ERR_ARG;
  AEW_LAUNCH(k5_extract, 1, 32, 0, (cudaStream_t)stream, d_fx, d_fy, baseline, d_xt, n_t, d_coords, d_meta);
  return (int)cudaGetLastError();
}

// ---- discrete sigma grids: separable tables (see k5_tables).  d_vals [n_vals <= 5] are the grid values,
// d_idx [n_cand][12] (uint8) the index of every candidate entry in that grid.
extern "C" int aew_recon_max_grid_values(void) { return K5_NV; }
extern "C" size_t aew_recon_tables_workspace_bytes(long long n_beats) {
  return (size_t)(k5_grid(n_beats) + 1) * K5_TABLE * sizeof(double) + 256;
}
extern "C" int aew_recon_sse_grid_f64(const double* d_fx, const double* d_fy, const double* d_baseline,
                                      const float* d_target, long long n_beats, const double* d_xt, int n_t,
                                      const double* d_vals, int n_vals, const unsigned char* d_idx, long long n_cand,
                                      double* d_sse, void* d_ws, size_t ws_bytes, void* stream) {
  if (!d_fx || !d_fy || !d_baseline || !d_target || !d_xt || !d_vals || !d_idx || !d_sse || !d_ws || n_beats < 0 ||
      n_cand < 1)
    return AEW_ERR_ARG;
  if (n_t < 2 || n_t > K5_NT || n_vals < 1 || n_vals > K5_NV) return AEW_ERR_UNSUPPORTED;
  if (ws_bytes < aew_recon_tables_workspace_bytes(n_beats)) return AEW_ERR_WORKSPACE;
  const int grid = k5_grid(n_beats);
  double* partial = (double*)(((size_t)d_ws + 127) & ~(size_t)127);
  double* tables = partial + (size_t)grid * K5_TABLE;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(partial, 0, (size_t)(grid + 1) * K5_TABLE * sizeof(double), st);
  AEW_LAUNCH(k5_tables, grid, K5_NT, 0, st, d_fx, d_fy, d_baseline, d_target, n_beats, d_xt, n_t, d_vals, n_vals, partial);
  AEW_LAUNCH(k5_tables_reduce, (K5_TABLE + 127) / 128, 128, 0, st, (const double*)partial, grid, tables);
  AEW_LAUNCH(k5_score, (unsigned)((n_cand + 127) / 128), 128, 0, st, (const double*)tables, d_idx, n_cand, d_sse);
  return (int)cudaGetLastError();
}
