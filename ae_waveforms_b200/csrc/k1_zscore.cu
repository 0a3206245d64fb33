// k1_zscore.cu -- whole-record NaN-aware z-score statistics and apply (K1).
//
// Replaces, per record of a ragged batch, the reference's
//     (data - np.nanmean(data)) / np.nanstd(data)
// (data_load_preprocess/filters.py:5-12, called at data_extract_vitaldb.py:64).
// mean and std are BIT-IDENTICAL to NumPy's: the sums follow NumPy's pairwise
// tree (aew_pairwise.cuh), NaN -> 0 in the sums, divisor = count of non-NaN.
//
// Data layout: one flat float64 array holding all records back to back,
// rec_off[n_rec+1] (int64) giving each record's range.  Work decomposition: a
// 256-thread block owns 32 consecutive leaf slots of one record (8 lanes per
// slot, 64-byte coalesced reads, up to 16 independent loads in flight per lane)
// and folds them 5 tree levels into one node; a second tiny kernel folds the
// nodes of each record to the root.  blk_start[n_rec+1] (int64, from
// aew_zscore_plan) maps block -> record.
// HBM traffic: 8 B/sample for the mean pass, 8 B/sample for the variance pass
// (+ 16 B/sample for a materialised apply; the filter kernel fuses the apply
// into its load instead).  Roofline: HBM.
#include "aew_pairwise.cuh"

namespace {

struct LoadSum {   // pass 1: NaN -> 0, count the non-NaN
  const double* x;
  int cnt;
  __device__ __forceinline__ double operator()(long long i, bool primary) {
    const double v = __ldg(x + i);
    const bool ok = !isnan(v);
    if (primary && ok) ++cnt;
    return ok ? v : 0.0;
  }
};
struct LoadDev2 {  // pass 2: squared deviation from the mean, NaN -> 0
  const double* x;
  double mean;
  int cnt;
  __device__ __forceinline__ double operator()(long long i, bool) {
    const double v = __ldg(x + i);
    if (isnan(v)) return 0.0;
    const double d = __dsub_rn(v, mean);
    return __dmul_rn(d, d);
  }
};

// Leaf sum of K1 with the loads of a lane issued back to back (up to 16 x 8 B in flight per lane).
// The kernel used to be bound by instruction issue, not by HBM (ncu: issue slots 75 % busy), so the common case is
// kept to a load and an addition per sample: the sums are formed WITHOUT looking for NaN -- a NaN poisons its
// partial sum, the xor steps spread it to all eight lanes -- and only a leaf whose plain sum is NaN is redone
// with the NaN -> 0 / counting arithmetic of np.nanmean / np.nanstd.  Same additions in the same order either way.
// the literal nan-aware leaf sum (values re-read: the fast path below does not keep them)
template <int PASS>
__device__ __noinline__ double k1_leaf_nan(const double* __restrict__ x, double mean, long long off, int size, int g,
                                           int& cnt) {
  double v[16];
  const int c8 = size >> 3;
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = (k < c8) ? __ldg(x + off + 8 * k + g) : 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const bool ok = !isnan(v[k]);
    if (PASS == 1) {
      cnt += (k < c8 && ok) ? 1 : 0;
      v[k] = ok ? v[k] : 0.0;
    } else {
      const double d = __dsub_rn(v[k], mean);
      v[k] = ok ? __dmul_rn(d, d) : 0.0;
    }
  }
  double r = v[0];
#pragma unroll
  for (int k = 1; k < 16; ++k)
    if (k < c8) r = __dadd_rn(r, v[k]);
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 1));
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 2));
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 4));
  int i = c8 << 3;
  if (size < 8) { r = 0.0; i = 0; }
  for (; i < size; ++i) {
    const double w = __ldg(x + off + i);
    const bool ok = !isnan(w);
    if (PASS == 1) {
      cnt += (g == 0 && ok) ? 1 : 0;
      r = __dadd_rn(r, ok ? w : 0.0);
    } else {
      const double d = __dsub_rn(w, mean);
      r = __dadd_rn(r, ok ? __dmul_rn(d, d) : 0.0);
    }
  }
  return r;
}

template <int PASS>
__device__ __forceinline__ double k1_leaf_sum8(const double* __restrict__ x, double mean, long long off, int size,
                                               int g, int& cnt) {
  double v[16];
  const int c8 = size >> 3;
  const double* xg = x + off + g;
  if (c8 >= 12) {                                       // (uniform over the 8 lanes; leaves of long records: 13 or 14)
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = __ldg(xg + 8 * k);
#pragma unroll
    for (int k = 12; k < 16; ++k) v[k] = (k < c8) ? __ldg(xg + 8 * k) : 0.0;
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = (k < c8) ? __ldg(xg + 8 * k) : 0.0;
  }
  if (PASS == 2) {
#pragma unroll
    for (int k = 0; k < 16; ++k) { const double d = __dsub_rn(v[k], mean); v[k] = __dmul_rn(d, d); }
  }
  double r = v[0];
  if (c8 >= 12) {
#pragma unroll
    for (int k = 1; k < 12; ++k) r = __dadd_rn(r, v[k]);
#pragma unroll
    for (int k = 12; k < 16; ++k) if (k < c8) r = __dadd_rn(r, v[k]);
  } else {
#pragma unroll
    for (int k = 1; k < 16; ++k) if (k < c8) r = __dadd_rn(r, v[k]);
  }
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 1));
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 2));
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 4));
  int i = c8 << 3;
  if (size < 8) { r = 0.0; i = 0; }
  for (; i < size; ++i) {
    const double w = __ldg(x + off + i);
    if (PASS == 1) r = __dadd_rn(r, w);
    else { const double d = __dsub_rn(w, mean); r = __dadd_rn(r, __dmul_rn(d, d)); }
  }
  // (the eight lanes of a leaf hold the same r; the four leaves of a warp go through the redo together because
  //  its exchanges involve the whole warp -- a leaf without NaN just obtains the same sum again)
  if (!__any_sync(AEW_FULL, isnan(r))) {
    if (PASS == 1) cnt += (g == 0) ? size : 0;
    return r;
  }
  return k1_leaf_nan<PASS>(x, mean, off, size, g, cnt);
}

// the record block b belongs to.  Records of equal length -- the usual batch -- put it at b n_rec / n_blocks: two
// independent loads confirm the guess; a binary search (log2 n_rec DEPENDENT loads in front of everything a block
// does) only where they do not
__device__ __forceinline__ long long find_record(const long long* __restrict__ start, long long n_rec,
                                                 long long b, long long n_blocks) {
  if (n_blocks > 0) {                                 // (start[] has n_rec + 1 entries then)
    long long g = b * n_rec / n_blocks;
    if (g > n_rec - 1) g = n_rec - 1;
    if (start[g] <= b && b < start[g + 1]) return g;
  }
  long long lo = 0, hi = n_rec;  // start[lo] <= b < start[hi]
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if (start[mid] <= b) lo = mid; else hi = mid;
  }
  return lo;
}

// PASS 1: nodes <- sums of x (NaN->0), cnt[r] += non-NaN count.
// PASS 2: nodes <- sums of (x-mean[r])^2.
template <int PASS>
__global__ void __launch_bounds__(256) k1_leaf(const double* __restrict__ x, const long long* __restrict__ rec_beg, const long long* __restrict__ rec_end,
                                               const long long* __restrict__ blk_start, long long n_rec,
                                               const double* __restrict__ mean, double* __restrict__ nodes,
                                               unsigned long long* __restrict__ cnt) {
  __shared__ double s_vals[32];
  __shared__ unsigned char s_valid[32];
  __shared__ long long s_rec, s_base, s_off, s_size;
  __shared__ int s_cnt, s_sub;
  const long long b = blockIdx.x;
  if (threadIdx.x == 0) {
    // the record, and the subtree of its NumPy tree that this block sums: walked once, not by every thread
    const long long r0 = find_record(blk_start, n_rec, b, (long long)gridDim.x);
    const long long base0 = rec_beg[r0], n0 = rec_end[r0] - base0;
    const int depth = pw_depth(n0);
    const int sub0 = depth < 5 ? depth : 5;
    long long o, sz;
    pw_node(n0, depth - sub0, (unsigned)(b - blk_start[r0]), o, sz);
    s_rec = r0; s_base = base0; s_off = o; s_size = sz; s_sub = sub0; s_cnt = 0;
  }
  __syncthreads();
  const long long r = s_rec;
  const long long base = s_base;
  const int sub = s_sub;
  const int group = threadIdx.x >> 3, g = threadIdx.x & 7;
  long long off = s_off, size = s_size;
  if (group < (1 << sub)) pw_walk(off, size, sub, (unsigned)group); else size = 0;
  int mycnt = 0;
  const double v = k1_leaf_sum8<PASS>(x + base, PASS == 2 ? mean[r] : 0.0, off, (int)size, g, mycnt);
  if (g == 0) { s_vals[group] = v; s_valid[group] = size > 0; }
  if (PASS == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mycnt += __shfl_xor_sync(AEW_FULL, mycnt, o);
    if ((threadIdx.x & 31) == 0 && mycnt) atomicAdd(&s_cnt, mycnt);
  }
  __syncthreads();
  if (threadIdx.x < 32) pw_warp_combine(s_vals, s_valid, sub, threadIdx.x);
  if (threadIdx.x == 0) {
    nodes[b] = s_vals[0];
    if (PASS == 1 && s_cnt) atomicAdd(cnt + r, (unsigned long long)s_cnt);
  }
}

// Fold the nodes of each record (2^(depth-5) of them) to the root.
// PASS 1 writes mean[r] = sum / cnt; PASS 2 writes sd[r] = sqrt(sum / cnt).
template <int PASS>
__global__ void __launch_bounds__(256) k1_combine(const long long* __restrict__ rec_beg, const long long* __restrict__ rec_end,
                                                  const long long* __restrict__ blk_start,
                                                  double* __restrict__ nodes, double* __restrict__ nodes2,
                                                  const unsigned long long* __restrict__ cnt,
                                                  double* __restrict__ out) {
  const long long r = blockIdx.x;
  const long long n = rec_end[r] - rec_beg[r];
  const int depth = pw_depth(n);
  const int dn = depth > 5 ? depth - 5 : 0;
  double* src = nodes + blk_start[r];
  double* dst = nodes2 + blk_start[r];
  for (int lvl = dn - 1; lvl >= 0; --lvl) {
    const int c = 1 << lvl;
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
      long long o, sa, sb;
      pw_node(n, lvl + 1, 2u * i, o, sa);
      pw_node(n, lvl + 1, 2u * i + 1u, o, sb);
      const double a = src[2 * i], bb = src[2 * i + 1];
      dst[i] = (sa > 0 && sb > 0) ? __dadd_rn(a, bb) : a;
    }
    __threadfence_block();
    __syncthreads();
    double* t = src; src = dst; dst = t;
  }
  if (threadIdx.x == 0) {
    const double q = __ddiv_rn(src[0], (double)cnt[r]);
    out[r] = PASS == 1 ? q : __dsqrt_rn(q);
  }
}

// ---- the statistics in ONE sweep (aew_zscore_stats_fast_f64): for callers that do not need NumPy's bits -- the
// tiled filter behind it is ~1e-9 from SciPy's output anyway -- 8 B/sample instead of 16.  A block takes the same
// range of its record as k1_leaf (a subtree of 32 leaves, ~3,500 samples) and reduces it to (n, mean, M2) around a
// shift of its own (its first sample: the sums of d = x - c and d^2 cancel nothing however far the record sits from
// zero); k1_fast_combine merges the triples of a record in a fixed order (Chan et al.): deterministic, and within a
// few ulp of np.nanmean / np.nanstd.  A NaN poisons the plain sums; only such a block repeats its range NaN-aware.
// (Measured and not kept, profiles/run_r2d_k1s.sh: PERSISTENT blocks that reduce unit i out of one shared-memory
// buffer while 8-byte cp.async copies of unit i + 1 fill another and a ninth warp looks up the range of unit i + 2 --
// the same bits in `nodes`, but 4.35 ms per 1,000 one-hour cases against 3.63: three blocks x one 28 KB unit in
// flight per SM is less than the 16 loads per thread x 1,024 threads this kernel keeps in its REGISTERS, and a
// block's iteration cannot be shorter than the loaded memory latency.)
// (Measured and not kept, profiles/run_r2d_k1bps.sh: five or six resident blocks per SM instead of the four that 62
// registers allow -- __launch_bounds__(256, 5 / 6) caps the kernel at 48 / 40 registers, the 16 loads in flight per
// thread spill, 5.0 / 5.1 ms per 1,000 one-hour cases against 3.6.)
__global__ void __launch_bounds__(256) k1_fast_leaf(const double* __restrict__ x, const long long* __restrict__ rec_beg,
                                                    const long long* __restrict__ rec_end,
                                                    const long long* __restrict__ blk_start, long long n_rec,
                                                    double* __restrict__ nodes) {
  __shared__ double s_red[3][8];
  __shared__ long long s_lo, s_n;
  __shared__ double s_c;
  __shared__ int s_redo;
  const long long b = blockIdx.x;
  if (threadIdx.x == 0) {
    const long long r0 = find_record(blk_start, n_rec, b, (long long)gridDim.x);
    const long long base0 = rec_beg[r0], n0 = rec_end[r0] - base0;
    const int depth = pw_depth(n0);
    const int sub0 = depth < 5 ? depth : 5;
    long long o, sz;
    pw_node(n0, depth - sub0, (unsigned)(b - blk_start[r0]), o, sz);
    s_lo = base0 + o; s_n = sz;
    s_redo = 0;
  }
  __syncthreads();
  const double* xb = x + s_lo;
  const int n = (int)s_n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double cntd = (double)n;
  // a block's range is 32 leaves of at most 128 samples: 16 per thread, every load issued before the first use.  The
  // shift is the block's first sample, i.e. thread 0's first value: it is handed round AFTER the loads have been
  // issued (read by thread 0 beforehand it was one more DRAM round trip in front of them)
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = (threadIdx.x + 256 * k < n) ? __ldg(xb + threadIdx.x + 256 * k) : 0.0;
  if (threadIdx.x == 0) {
    double c0 = n > 0 ? v[0] : 0.0;
    if (isnan(c0) && n > 0) c0 = xb[n - 1];
    if (isnan(c0) && n > 0) c0 = xb[n / 2];
    s_c = isnan(c0) ? 0.0 : c0;
  }
  __syncthreads();
  const double c = s_c;
#pragma unroll 1
  for (int attempt = 0; attempt < 2; ++attempt) {
    double s1a = 0.0, s1b = 0.0, s2a = 0.0, s2b = 0.0, na = 0.0;
    const bool aware = attempt == 1;
    // (a sample that is not there counts as the shift: d = 0)
    {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const bool there = threadIdx.x + 256 * k < n;
        const bool ok = there && (!aware || !isnan(v[k]));
        const double d = ok ? v[k] - c : 0.0;
        if (aware) na += (ok && there) ? 1.0 : 0.0;
        if (k & 1) { s1b += d; s2b = __fma_rn(d, d, s2b); } else { s1a += d; s2a = __fma_rn(d, d, s2a); }
      }
    }
    for (int i = threadIdx.x + 16 * 256; i < n; i += 256) {   // (never, with the plan of aew_zscore_plan)
      const double vi = __ldg(xb + i);
      const bool ok = !aware || !isnan(vi);
      const double d = ok ? vi - c : 0.0;
      if (aware) na += ok ? 1.0 : 0.0;
      s1a += d; s2a = __fma_rn(d, d, s2a);
    }
    double s1 = s1a + s1b, s2 = s2a + s2b;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(AEW_FULL, s1, o);
      s2 += __shfl_xor_sync(AEW_FULL, s2, o);
      na += __shfl_xor_sync(AEW_FULL, na, o);
    }
    __syncthreads();                                      // (second attempt: the first one's values have been read)
    if (lane == 0) { s_red[0][warp] = s1; s_red[1][warp] = s2; s_red[2][warp] = na; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double S1 = 0.0, S2 = 0.0, N = 0.0;
      for (int k = 0; k < 8; ++k) { S1 += s_red[0][k]; S2 += s_red[1][k]; N += s_red[2][k]; }
      if (!aware && (isnan(S1) || isnan(S2))) {
        s_redo = 1;
      } else {
        if (!aware) N = cntd;
        const double md = N > 0.0 ? S1 / N : 0.0;
        nodes[3 * b] = N;
        nodes[3 * b + 1] = c + md;
        nodes[3 * b + 2] = N > 0.0 ? S2 - S1 * md : 0.0;
        s_redo = 0;
      }
    }
    __syncthreads();
    if (!s_redo) break;
  }
}

// one block per record: n = sum n_b, mean = sum n_b mean_b / n, M2 = sum (M2_b + n_b (mean_b - mean)^2), each sum a
// fixed tree over the record's blocks
__device__ __forceinline__ double k1_block_sum256(double v, double* s_w) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(AEW_FULL, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int k = 0; k < 8; ++k) t += s_w[k];
  return t;
}
__global__ void __launch_bounds__(256) k1_fast_combine(const long long* __restrict__ blk_start,
                                                       const double* __restrict__ nodes, double* __restrict__ mean,
                                                       double* __restrict__ sd, long long* __restrict__ cnt) {
  __shared__ double s_w[8];
  const long long r = blockIdx.x;
  const long long b0 = blk_start[r], b1 = blk_start[r + 1];
  double n = 0.0, nm = 0.0;
  for (long long b = b0 + threadIdx.x; b < b1; b += 256) { n += nodes[3 * b]; nm += nodes[3 * b] * nodes[3 * b + 1]; }
  const double N = k1_block_sum256(n, s_w);
  const double NM = k1_block_sum256(nm, s_w);
  const double m = NM / N;                                 // (an all-NaN record: 0 / 0 = NaN, as np.nanmean warns and gives)
  double m2 = 0.0;
  for (long long b = b0 + threadIdx.x; b < b1; b += 256) {
    const double d = nodes[3 * b + 1] - m;
    m2 += nodes[3 * b + 2] + nodes[3 * b] * d * d;
  }
  const double M2 = k1_block_sum256(m2, s_w);
  if (threadIdx.x == 0) {
    mean[r] = m;
    sd[r] = sqrt(M2 / N);
    if (cnt) cnt[r] = (long long)N;
  }
}

// out[i] = (x[i] - mean[r]) / sd[r] for i inside record r; NaN between records
__global__ void __launch_bounds__(256) k1_apply(const double* __restrict__ x, const long long* __restrict__ rec_beg,
                                                const long long* __restrict__ rec_end, long long n_rec,
                                                const double* __restrict__ mean, const double* __restrict__ sd,
                                                double* __restrict__ out, long long lo, long long hi) {
  const long long per_thread = 4;
  long long i = lo + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * per_thread;
  if (i >= hi) return;
  long long r = find_record(rec_beg, n_rec, i, 0);
  for (int k = 0; k < per_thread && i < hi; ++k, ++i) {
    while (r + 1 < n_rec && i >= rec_beg[r + 1]) ++r;
    out[i] = (i >= rec_beg[r] && i < rec_end[r]) ? __ddiv_rn(__dsub_rn(x[i], mean[r]), sd[r])
                                                 : __longlong_as_double(0x7ff8000000000000ll);
  }
}

// first / one-past-last non-NaN sample of every record (data_extract_vitaldb.py:53-57):
// trimmed[2r] = begin, trimmed[2r+1] = end (begin == end when the record is all NaN)
__global__ void __launch_bounds__(256) k1_trim(const double* __restrict__ x, const long long* __restrict__ rec_beg,
                                               const long long* __restrict__ rec_end,
                                               long long* __restrict__ trimmed) {
  const long long r = blockIdx.x;
  const long long b = rec_beg[r], e = rec_end[r];
  __shared__ long long s_first, s_last;
  if (threadIdx.x == 0) { s_first = e; s_last = b - 1; }
  __syncthreads();
  for (long long base = b; base < e; base += blockDim.x) {       // from the front
    const long long i = base + threadIdx.x;
    const int ok = (i < e) && !isnan(x[i]);
    if (ok) atomicMin(&s_first, i);
    if (__syncthreads_or(ok)) break;
  }
  for (long long top = e; top > b; top -= blockDim.x) {          // from the back
    const long long i = top - 1 - threadIdx.x;
    const int ok = (i >= b) && !isnan(x[i]);
    if (ok) atomicMax(&s_last, i);
    if (__syncthreads_or(ok)) break;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const bool any = s_first < e;
    trimmed[2 * r] = any ? s_first : b;
    trimmed[2 * r + 1] = any ? s_last + 1 : b;
  }
}

}  // namespace

// ---------------------------------------------------------------- host side
extern "C" long long aew_zscore_plan(const long long* h_rec_beg, const long long* h_rec_end, long long n_rec,
                                     long long* h_blk_start) {
  long long acc = 0;
  for (long long r = 0; r < n_rec; ++r) {
    if (h_blk_start) h_blk_start[r] = acc;
    const int d = pw_depth(h_rec_end[r] - h_rec_beg[r]);
    acc += 1ll << (d > 5 ? d - 5 : 0);
  }
  if (h_blk_start) h_blk_start[n_rec] = acc;
  return acc;  // number of leaf blocks == number of tree nodes
}

extern "C" size_t aew_zscore_workspace_bytes(long long n_rec, long long n_blocks) {
  return (size_t)(3 * n_blocks) * sizeof(double) + (size_t)n_rec * sizeof(unsigned long long) + 64;
}

extern "C" int aew_trim_nan_edges_f64(const double* x, const long long* d_rec_beg, const long long* d_rec_end,
                                      long long n_rec, long long* d_trimmed, void* stream) {
  if (!x || !d_rec_beg || !d_rec_end || !d_trimmed || n_rec < 0) return AEW_ERR_ARG;
  if (n_rec == 0) return AEW_OK;
  if (n_rec > 0x7fffffffll) return AEW_ERR_UNSUPPORTED;
  AEW_LAUNCH(k1_trim, (unsigned)n_rec, 256, 0, (cudaStream_t)stream, x, d_rec_beg, d_rec_end, d_trimmed);
  return (int)cudaGetLastError();
}

extern "C" int aew_zscore_stats_f64(const double* x, const long long* d_rec_beg, const long long* d_rec_end,
                                    const long long* d_blk_start, long long n_rec, long long n_blocks, double* mean,
                                    double* sd, long long* count, void* ws, size_t ws_bytes, void* stream) {
  if (!x || !d_rec_beg || !d_rec_end || !d_blk_start || !mean || !sd || !ws || n_rec < 0) return AEW_ERR_ARG;
  if (n_rec == 0) return AEW_OK;
  if (ws_bytes < aew_zscore_workspace_bytes(n_rec, n_blocks)) return AEW_ERR_WORKSPACE;
  if (n_blocks > 0x7fffffffll || n_rec > 0x7fffffffll) return AEW_ERR_UNSUPPORTED;
  double* nodes = (double*)ws;
  double* nodes2 = nodes + n_blocks;
  unsigned long long* cnt = count ? (unsigned long long*)count : (unsigned long long*)(nodes2 + n_blocks);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(cnt, 0, (size_t)n_rec * sizeof(unsigned long long), st);
  AEW_LAUNCH(k1_leaf<1>, (unsigned)n_blocks, 256, 0, st, x, d_rec_beg, d_rec_end, d_blk_start, n_rec,
             (const double*)nullptr, nodes, cnt);
  AEW_LAUNCH(k1_combine<1>, (unsigned)n_rec, 256, 0, st, d_rec_beg, d_rec_end, d_blk_start, nodes, nodes2,
             (const unsigned long long*)cnt, mean);
  AEW_LAUNCH(k1_leaf<2>, (unsigned)n_blocks, 256, 0, st, x, d_rec_beg, d_rec_end, d_blk_start, n_rec,
             (const double*)mean, nodes, cnt);
  AEW_LAUNCH(k1_combine<2>, (unsigned)n_rec, 256, 0, st, d_rec_beg, d_rec_end, d_blk_start, nodes, nodes2,
             (const unsigned long long*)cnt, sd);
  return (int)cudaGetLastError();
}

// the same three outputs from one sweep over the samples, within a few ulp of NumPy's (not bit-identical)
extern "C" int aew_zscore_stats_fast_f64(const double* x, const long long* d_rec_beg, const long long* d_rec_end,
                                         const long long* d_blk_start, long long n_rec, long long n_blocks,
                                         double* mean, double* sd, long long* count, void* ws, size_t ws_bytes,
                                         void* stream) {
  if (!x || !d_rec_beg || !d_rec_end || !d_blk_start || !mean || !sd || !ws || n_rec < 0) return AEW_ERR_ARG;
  if (n_rec == 0) return AEW_OK;
  if (ws_bytes < aew_zscore_workspace_bytes(n_rec, n_blocks)) return AEW_ERR_WORKSPACE;
  if (n_blocks > 0x7fffffffll || n_rec > 0x7fffffffll) return AEW_ERR_UNSUPPORTED;
  double* nodes = (double*)ws;
  cudaStream_t st = (cudaStream_t)stream;
  AEW_LAUNCH(k1_fast_leaf, (unsigned)n_blocks, 256, 0, st, x, d_rec_beg, d_rec_end, d_blk_start, n_rec, nodes);
  AEW_LAUNCH(k1_fast_combine, (unsigned)n_rec, 256, 0, st, d_blk_start, (const double*)nodes, mean, sd, count);
  return (int)cudaGetLastError();
}

// applies over flat positions [lo, hi); positions outside every record get NaN
extern "C" int aew_zscore_apply_f64(const double* x, const long long* d_rec_beg, const long long* d_rec_end,
                                    long long n_rec, const double* mean, const double* sd, double* out,
                                    long long lo, long long hi, void* stream) {
  if (!x || !d_rec_beg || !d_rec_end || !mean || !sd || !out || n_rec < 1 || hi < lo) return AEW_ERR_ARG;
  if (hi == lo) return AEW_OK;
  const long long threads = (hi - lo + 3) / 4;
  const long long blocks = (threads + 255) / 256;
  if (blocks > 0x7fffffffll) return AEW_ERR_UNSUPPORTED;
  AEW_LAUNCH(k1_apply, (unsigned)blocks, 256, 0, (cudaStream_t)stream, x, d_rec_beg, d_rec_end, n_rec, mean, sd, out,
             lo, hi);
  return (int)cudaGetLastError();
}
