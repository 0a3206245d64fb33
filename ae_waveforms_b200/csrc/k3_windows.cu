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
#include <cstdlib>
#include "aew_pairwise.cuh"
#include "aew_math.cuh"
#include "aew_bulk.cuh"
#include "../../include/aewave.h"

#define K3_NT 512                         // two blocks per SM = 32 warps (64 registers per thread): the kernel is bound by
                                          // issue latency, not by a pipe (profiles/r1_k3_screen.txt: 16 warps, 45 % issue)
#define K3_NW (K3_NT / 32)
#define K3_ITEMS 22                       // extended positions per thread: ONE strip of the whole window (even: the
                                          // strip's own accesses are 16-byte pairs, conflict-free at a stride of 22)
#define K3_WCAP (K3_NT * K3_ITEMS)        // 11,264 positions: window + the two box filters' halo
#define K3_MAXRUN 2048                    // QRS runs per window the tables hold (a 20 s window has ~25)
#define K3_RING 512                       // ring of smoothgrad values in the sequential evaluation (> avg_k + 1)
#define K3_MAXL 10240                     // samples per window at most (80 KB of shared memory)
#define K3_QWS (K3_MAXL / 32 + 2)         // words of a window's indicator bit array
#define K3_MAXP 72                        // peaks are > 150 samples apart -> at most 71 in 10752
#define K3_MAXS 512                       // beat section int(0.64 fs) at most (320 at 500 Hz; median_sample.py:20-21)
#define K3_NONE 0xffffu
#define K3_PREFETCH_AHEAD 296             // one wave of blocks (two blocks per SM)

namespace {

struct K3Smem {            // shared memory of the window kernel k3_windows
  double* W;            // [K3_WCAP] the window buffer: z-scored samples at W + H, then -- in place -- the prefix sums of
                        //           |gradient|, then those of smoothgrad
  double* pw_vals;      // [2][128] leaf sums, double-buffered
  double* wtot;         // [2][16]
  double* scal;         // [16] scalars
  unsigned* Qw;         // [K3_QWS] QRS indicator, one bit per sample
  double* ring;         // [K3_RING] smoothgrad ring of the sequential evaluation
  unsigned long long* mbar;   // mbarrier of the bulk copy that brings the window in
  unsigned char* pw_valid;  // [2][128]
};

// what the window kernel hands to the beat kernel, per window (device workspace of the call)
struct K3Ws {
  unsigned* qbits;      // [n_win][K3_QWS] indicator bits
  double* sumz;         // [n_win] NumPy-order sum of the z-scored window (np.std's mean)
  double* sd;           // [n_win] np.std of the z-scored window where the window kernel formed it (lazy_sd == 0)
  double* mean;         // [n_win] the window's own mean and sd (its z-score): the median kernel forms the z-scores of
  double* sd0;          //         the two middle values of a column from them; with fast_stats the beat kernel too
  unsigned char* route; // [n_win] which median kernel takes the window: 0 none (removed), 1 k3_median_keys<32>,
                        //         2 k3_median_keys<64>, 3 k3_median (also set by a keys kernel that hands a window back)
};

// The NumPy-tree leaves of a length-L vector, tabulated by the host per launch (they depend on L alone), and the
// leaf each group of four threads serves.  Leaf offsets are multiples of 8 samples = 16 banks: two leaves whose
// offsets agree modulo 16 samples collide when the same quarter-warp reads them (about half of the neighbouring
// pairs at L = 10,000: 21 % of the kernel's shared-memory wavefronts in profiles/r2_ncu_k3.txt), so groups 2q and
// 2q + 1 are given one leaf of either residue.
struct K3Leaves {
  unsigned short off[128];
  unsigned char size[128];
  unsigned char slot[128];    // leaf served by thread group g = tid >> 2
  int depth, all;             // all: every one of the 2^depth = 128 slots holds samples
};

// first sample's offset inside W: the halo of the two box filters in front of it (even: 16-byte pairs stay aligned)
__host__ __device__ inline int k3_halo(int ks, int ka) { return (ka / 2 + ks / 2 + 2) & ~1; }
// extended positions the indicator needs: sample j sits at position j + H, the averaging box of the last sample
// ends ka - 1 - ka/2 samples behind it
__host__ __device__ inline int k3_ext_len(int L, int ks, int ka) { return L + k3_halo(ks, ka) + ka - ka / 2 - 1; }

// Everything is of fixed size, at compile-time offsets from the base of the dynamic shared memory (with
// L-dependent offsets the compiler re-derived these pointers at their uses under the 64-register cap: 7.6 % of all
// executed instructions in profiles/r2_k3_512_before_diet.txt); the window buffer comes last.
__host__ __device__ inline size_t k3_smem_layout(unsigned char* base, K3Smem* s) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 15) & ~(size_t)15; return base ? base + at : (unsigned char*)0; };
  K3Smem t;
  t.pw_vals = (double*)take(sizeof(double) * 256);
  t.wtot = (double*)take(sizeof(double) * 2 * K3_NW);
  t.scal = (double*)take(sizeof(double) * 16);
  t.pw_valid = (unsigned char*)take(256);
  t.mbar = (unsigned long long*)take(16);
  t.Qw = (unsigned*)take(sizeof(unsigned) * K3_QWS);
  t.ring = (double*)take(sizeof(double) * K3_RING);
  t.W = (double*)take(sizeof(double) * K3_WCAP);
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

#if defined(AEW_EMUL) || !defined(AEW_K3_TICKS)     // (an instrumented build: nvcc -DAEW_K3_TICKS; the counters cost registers)
#define K3_TICK(slot)
#else
#define K3_TICK(slot)                                                              \
  if (dbg && tid == 0) {                                                           \
    const long long t_now = clock64();                                             \
    atomicAdd(dbg + (slot), (unsigned long long)(t_now - t_last));                 \
    t_last = t_now;                                                                \
  }
#endif

// ---- passes over the window in NumPy's summation order -------------------------------------------------------
// Thread t owns lane j = t & 3 of leaf t >> 2 of the NumPy tree (128 leaves of 72..88 samples for a 10,000-sample
// window, so one round of the 512 threads covers the window).  Lane j carries the stride-8 partial sums r_j and
// r_{j+4}; ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) falls out of two xor steps and one local addition, the leaf's
// tail is added sequentially.  A pass is described by a functor: load(i) fetches element i, apply(v, i, primary)
// returns the term to accumulate and may have side effects (stores) when `primary` -- true on exactly one lane per
// element.  Work that used to be separate sweeps rides along: the HBM load + NaN screen with the first sum, the
// z-score transform with the sum of the transformed window.
// The kernel is bound by INSTRUCTION FETCH (profiles/icache_bench.cu: straight-line code that does not fit the
// 32 KB L1.5 instruction cache streams from L2 at about 5.5 cycles per instruction per SM, and every window runs
// through the whole kernel once), so everything used more than once is one non-inlined function.
enum { K3_PASS_PLAIN = 0, K3_PASS_DEV2 = 1, K3_PASS_ZSCORE = 2, K3_PASS_ZDEV2 = 3 };
struct LeafOut { double r; int flag; };
struct SumPlain { const double* a; __device__ __forceinline__ double operator()(int i, bool) { return a[i]; } };

// One sweep over the thread's share of its leaf.  PLAIN: sum of the window, flag = saw a NaN
// (remove_samples.py:31-33).  DEV2: sum of (x - m)^2.  ZSCORE: z = (x - m) / sd stored to D, sum of the z,
// flag = produced a non-finite value (constant window: 0/0).  ZDEV2: sum of (z - m2)^2 with z formed on the fly
// (np.std of the z-scored window where it is wanted for every window).
__device__ __forceinline__ double leaf_term(int mode, double v, double m, const ExactDiv& dv, double m2, int& flag) {
  if (mode == K3_PASS_PLAIN) { flag |= isnan(v) ? 1 : 0; return v; }
  const double d = __dsub_rn(v, m);
  if (mode == K3_PASS_DEV2) return __dmul_rn(d, d);
  const double z = div_exact(d, dv);
  if (mode == K3_PASS_ZDEV2) { const double e = __dsub_rn(z, m2); return __dmul_rn(e, e); }
  flag |= (isnan(z) || isinf(z)) ? 1 : 0;
  return z;
}
// The general form: every element tested (NaN / non-finite result), IEEE division through div_exact, 8-byte
// accesses.  S: the samples (shared or global memory); D: where ZSCORE stores (may be S itself: every element is
// read before it is written).
__device__ __noinline__ LeafOut leaf_pass_checked(int mode, const double* S, double* D, double m, double sd, double m2,
                                                  int off, int size, int j) {
  const ExactDiv dv = make_exact_div(mode >= K3_PASS_ZSCORE ? sd : 1.0);
  const int c8 = size >> 3;
  int flag = 0;
  double r0 = 0.0, r1 = 0.0;
  const double* Sj = S + off + 2 * j;
  double* Dj = D + off + 2 * j;
#pragma unroll 1
  for (int k = 0; k < c8; ++k) {
    const double tx = leaf_term(mode, Sj[8 * k], m, dv, m2, flag);
    const double ty = leaf_term(mode, Sj[8 * k + 1], m, dv, m2, flag);
    if (mode == K3_PASS_ZSCORE) { Dj[8 * k] = tx; Dj[8 * k + 1] = ty; }
    r0 = k == 0 ? tx : __dadd_rn(r0, tx);
    r1 = k == 0 ? ty : __dadd_rn(r1, ty);
  }
  double r = __dadd_rn(r0, r1);
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 1));
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 2));
  int i = c8 << 3;
  if (size < 8) { r = 0.0; i = 0; }
  // the tail is added by all four lanes alike; each element is read by all before lane 0 may overwrite it
  int left = size - i;
  left = __reduce_max_sync(AEW_FULL, (unsigned)(left < 0 ? 0 : left));   // trip count uniform over the warp
#pragma unroll 1
  for (int u = 0; u < left; ++u) {
    const bool in = i + u < size;
    const double v = in ? S[off + i + u] : 0.0;
    __syncwarp();
    if (in) {
      const double t = leaf_term(mode, v, m, dv, m2, flag);
      if (mode == K3_PASS_ZSCORE && j == 0) D[off + i + u] = t;
      r = __dadd_rn(r, t);
    }
  }
  LeafOut o;
  o.r = r; o.flag = flag;
  return o;
}

// |q| outside [2^-300, 2^300]: zero, denormal-ish, huge, infinite or NaN (one shift, one add, one compare)
__device__ __forceinline__ unsigned odd_exponent(double q) {
  const unsigned e = ((unsigned)__double2hiint(q) >> 20) & 0x7ffu;
  return (e - 723u) > 600u ? 1u : 0u;
}

// The sweep proper over the window A in shared memory.  Lane j carries the stride-8 partial sums r_{2j} and
// r_{2j+1}: one 16-byte access per round (the leaf offsets are multiples of 8 samples), and
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is a local addition and two xor steps.  The common case runs WITHOUT
// per-element tests or branches (they were more instructions than the arithmetic:
// profiles/r2_k3_512_before_diet.txt): a NaN poisons the plain sum, and the Markstein division is applied to every
// element while an integer test on the first quotient's exponent collects the ones it must not be trusted with
// (zero, tiny, huge, non-finite).  Only a warp that saw such a thing repeats its leaves in the checked form above --
// same operations on the same operands in the same order, so the bits do not depend on the path.  ZSCORE leaves A
// as it is and stores the z-scores to D (the parked copy in global memory; d16: 16-byte stores allowed).
__device__ __noinline__ LeafOut leaf_pass(int mode, const double* A, double* D, bool d16, double m, double sd, int off,
                                          int size, int j) {
  const int c8 = size >> 3;
  double r0 = 0.0, r1 = 0.0;
  const double* Aj = A + off + 2 * j;
  unsigned odd = 0;
  if (mode == K3_PASS_PLAIN) {
    if (c8 > 0) { const double2 v = *reinterpret_cast<const double2*>(Aj); r0 = v.x; r1 = v.y; }
#pragma unroll 3
    for (int k = 1; k < c8; ++k) {
      const double2 v = *reinterpret_cast<const double2*>(Aj + 8 * k);
      r0 = __dadd_rn(r0, v.x);
      r1 = __dadd_rn(r1, v.y);
    }
  } else if (mode == K3_PASS_DEV2) {
    if (c8 > 0) {
      const double2 v = *reinterpret_cast<const double2*>(Aj);
      const double dx = __dsub_rn(v.x, m), dy = __dsub_rn(v.y, m);
      r0 = __dmul_rn(dx, dx); r1 = __dmul_rn(dy, dy);
    }
#pragma unroll 3
    for (int k = 1; k < c8; ++k) {
      const double2 v = *reinterpret_cast<const double2*>(Aj + 8 * k);
      const double dx = __dsub_rn(v.x, m), dy = __dsub_rn(v.y, m);
      r0 = __dadd_rn(r0, __dmul_rn(dx, dx));
      r1 = __dadd_rn(r1, __dmul_rn(dy, dy));
    }
  } else {
    const ExactDiv dv = make_exact_div(sd);
    odd = dv.fast ? 0u : 1u;
    if (dv.fast) {
      double* Dj = D + off + 2 * j;
#pragma unroll 2
      for (int k = 0; k < c8; ++k) {
        const double2 v = *reinterpret_cast<const double2*>(Aj + 8 * k);
        const double ax = __dsub_rn(v.x, m), ay = __dsub_rn(v.y, m);
        const double px = __dmul_rn(ax, dv.r), py = __dmul_rn(ay, dv.r);
        odd |= odd_exponent(px) | odd_exponent(py);
        const double qx = __fma_rn(__fma_rn(-px, dv.b, ax), dv.r, px), qy = __fma_rn(__fma_rn(-py, dv.b, ay), dv.r, py);
        double2 t;
        t.x = __fma_rn(__fma_rn(-qx, dv.b, ax), dv.r, qx);
        t.y = __fma_rn(__fma_rn(-qy, dv.b, ay), dv.r, qy);
        if (d16) *reinterpret_cast<double2*>(Dj + 8 * k) = t;
        else { Dj[8 * k] = t.x; Dj[8 * k + 1] = t.y; }
        r0 = k == 0 ? t.x : __dadd_rn(r0, t.x);
        r1 = k == 0 ? t.y : __dadd_rn(r1, t.y);
      }
    }
  }
  double r = __dadd_rn(r0, r1);
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 1));
  r = __dadd_rn(r, __shfl_xor_sync(AEW_FULL, r, 2));
  if (mode != K3_PASS_ZSCORE) odd = (size >= 8 && odd_exponent(r)) ? 1u : 0u;   // NaN / overflow shows in the sum
  if ((size & 7) != 0 || size < 8) odd = 1u;             // leaves with a tail (not at L = 10,000) take the general form
  if (__any_sync(AEW_FULL, odd != 0)) {
    // (a zero sum or quotient is "odd" too: harmless, the checked form gives the same bits)
    return leaf_pass_checked(mode, A, D, m, sd, 0.0, off, size, j);
  }
  LeafOut o;
  o.r = r; o.flag = 0;
  return o;
}

// Combine the leaf values up the tree: thread t has served leaf t >> 2 (512 threads = 128 leaves x 4 lanes).
// depth 7: every warp folds the 128 leaves in registers (two local levels, five xor steps) -- an invalid (empty)
// right node leaves its sibling untouched, exactly like pw_warp_combine -- so one barrier per sum is enough.  Two
// leaf buffers alternate (the caller flips `buf`): a slow warp may still be folding the previous sum while the next
// one's leaves are stored.
__device__ __noinline__ double block_combine(double ra, int size_a, int slot, int j, int depth,
                                             double* vals, unsigned char* valid, bool all_leaves) {
  const int nslots = 1 << depth;
  if (j == 0 && slot < nslots) { vals[slot] = ra; valid[slot] = size_a > 0; }
  __syncthreads();
  if (depth == 7 && all_leaves) {                         // every one of the 128 leaves exists: plain pairwise adds
    const int lane = threadIdx.x & 31;
    const double2 a01 = *reinterpret_cast<const double2*>(vals + 4 * lane);
    const double2 a23 = *reinterpret_cast<const double2*>(vals + 4 * lane + 2);
    double m = __dadd_rn(__dadd_rn(a01.x, a01.y), __dadd_rn(a23.x, a23.y));
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) m = __dadd_rn(m, __shfl_xor_sync(AEW_FULL, m, o));   // (IEEE + commutes)
    return m;
  }
  if (depth == 7) {
    const int lane = threadIdx.x & 31;
    const double a0 = vals[4 * lane], a1 = vals[4 * lane + 1], a2 = vals[4 * lane + 2], a3 = vals[4 * lane + 3];
    const int v0 = valid[4 * lane], v1 = valid[4 * lane + 1], v2 = valid[4 * lane + 2], v3 = valid[4 * lane + 3];
    const double n01 = (v0 && v1) ? __dadd_rn(a0, a1) : a0;
    const double n23 = (v2 && v3) ? __dadd_rn(a2, a3) : a2;
    double m = (v0 && v2) ? __dadd_rn(n01, n23) : n01;
    int mv = v0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double other = __shfl_xor_sync(AEW_FULL, m, o);
      const int ov = __shfl_xor_sync(AEW_FULL, mv, o);
      if ((lane & o) == 0) { m = (mv && ov) ? __dadd_rn(m, other) : m; }
      else { m = (ov && mv) ? __dadd_rn(other, m) : other; mv = ov; }
    }
    return m;
  }
  if (threadIdx.x < 32) pw_warp_combine(vals, valid, depth, threadIdx.x);
  __syncthreads();
  return vals[0];
}

// IEEE division / square root: ~40 instructions each when inlined, so one copy
__device__ __noinline__ double k3_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __noinline__ double k3_sqrt(double a) { return __dsqrt_rn(a); }

// exclusive block scan of one double per thread (fast path; not order-sensitive).  ONE barrier: consecutive scans
// alternate between two scratch rows (the caller flips `row`), so the warp totals of one scan cannot be
// overwritten before every warp has read them.
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
  double t = lane < K3_NW ? wtot[lane] : 0.0;            // every warp scans the K3_NW warp totals for itself
#pragma unroll
  for (int d = 1; d < K3_NW; d <<= 1) {
    const double o = __shfl_up_sync(AEW_FULL, t, d);
    if (lane >= d) t += o;
  }
  const double before = __shfl_sync(AEW_FULL, t, warp > 0 ? warp - 1 : 0);
  return (warp > 0 ? before : 0.0) + (incl - v);
}

// order-preserving 64-bit key of a (non-NaN) double: unsigned comparison of keys == comparison of the values
__device__ __forceinline__ unsigned long long dkey(double x) {
  const long long b = __double_as_longlong(x);
  return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  const long long t = (long long)k;
  return __longlong_as_double(t ^ ((~t >> 63) | (long long)0x8000000000000000ull));
}
// minimum over the warp of per-lane keys (REDUX on the high words, then on the low words of the lanes that tie)
__device__ __forceinline__ unsigned long long warp_min_key_raw(unsigned long long k) {
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mh = __reduce_min_sync(AEW_FULL, hi);
  const unsigned ml = __reduce_min_sync(AEW_FULL, hi == mh ? lo : 0xffffffffu);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ double warp_min_key(unsigned long long k) { return dkey_inv(warp_min_key_raw(k)); }

// ---------------------------------------------------------------------- medians
__device__ __forceinline__ void cex(double& a, double& b) {          // a <- min, b <- max (DSETP + 4 FSEL)
  const bool s = a > b;
  const double lo = s ? b : a, hi = s ? a : b;
  a = lo; b = hi;
}

// Batcher odd-even merge sort of N (power of two) values held in registers; the
// recursion is unfolded by templates so that every index is a compile-time constant
// (191 compare-exchanges for N = 32).  [LO, HI] inclusive.
__device__ __forceinline__ void cex(unsigned& a, unsigned& b) {      // integer keys: IMNMX x 2
  const unsigned lo = a < b ? a : b, hi = a < b ? b : a;
  a = lo; b = hi;
}
template <int LO, int HI, int R>
struct OddEvenMerge {
  template <class T, int N>
  static __device__ __forceinline__ void apply(T (&v)[N]) {
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
  template <class T, int N>
  static __device__ __forceinline__ void apply(T (&v)[N]) {
    if constexpr (HI - LO >= 1) {
      constexpr int MID = LO + (HI - LO) / 2;
      OddEvenSort<LO, MID>::apply(v);
      OddEvenSort<MID + 1, HI>::apply(v);
      OddEvenMerge<LO, HI, 1>::apply(v);
    }
  }
};
template <class T, int N>
__device__ __forceinline__ void sort_network(T (&v)[N]) { OddEvenSort<0, N - 1>::apply(v); }

// value of beat b at column c (median_sample.py:30-41): short slices are LEFT-aligned
// in a NaN row; +inf stands for the NaN padding
// where the samples of a window come from: as stored, or z-scored on the fly with the window's own mean / sd --
// the same two roundings as the in-place transform of the screen kernel, so the bits agree
struct BeatSrc {
  const double* z;
  double mean;
  ExactDiv dv;
  int zs;
  __device__ __forceinline__ double at(int idx) const {
    const double x = z[idx];
    return zs ? div_exact(__dsub_rn(x, mean), dv) : x;
  }
};
__device__ __forceinline__ double beat_value(const BeatSrc& z, int p, int c, int L, int pre, int post, int& valid) {
  const int s = p - pre < 0 ? 0 : p - pre;
  const int e = p + post > L ? L : p + post;
  const int idx = s + c;
  if (idx < e) { ++valid; return z.at(idx); }
  return __longlong_as_double(0x7ff0000000000000ll);
}

// One column of <= 64 beats.  n <= 32: ONE thread sorts the column's values with the 32-input network.
// 33..64 beats: a PAIR of adjacent lanes sorts 32 values each, a bitonic split across the pair leaves the 32
// smallest of the 64 with the even lane and the 32 largest with the odd lane (min / max of v[i] against the
// partner's v[31 - i]), and one 5-stage bitonic merge orders each half; 34 lane exchanges per column instead of
// the hundreds a warp-wide sort needs.  The median is (lo + hi) / 2 with lo == hi for odd counts, exactly as
// np.ma.median evaluates it; an all-NaN column gives NaN.  `paired` is uniform over the block; inactive
// threads run along (their lanes take part in the exchanges) and return garbage.
__device__ __forceinline__ double column_median_64(const BeatSrc& z, const int* __restrict__ peaks,
                                                   int n, int c, int half, bool active, bool paired, int L, int pre,
                                                   int post) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double v[32];
  // all loads first (no control flow between them: from global memory each one is a full round trip), then
  // the optional z-score of the values that exist
  unsigned okmask = 0;
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    const int bb = 32 * half + b;
    const bool have = active && bb < n;
    const int pk = have ? peaks[bb] : 0;
    const int s0 = pk - pre < 0 ? 0 : pk - pre;
    const int e0 = pk + post > L ? L : pk + post;
    const bool ok = have && s0 + c < e0;                 // short slices are left-aligned, NaN-padded on the right
    v[b] = ok ? z.z[s0 + c] : inf;
    okmask |= (unsigned)ok << b;
  }
  // z.zs: the samples are unscaled and the median is wanted of their z-scores (x - mean) / sd.  The z-score is a
  // non-decreasing map (rounding is monotonic, sd > 0), so the k-th smallest z-score is the z-score of the k-th
  // smallest sample: the column is sorted as it is and only the one or two middle values are transformed -- the
  // same two roundings on the same operands as np.median of the z-scored rows sees, two divisions per column
  // instead of one per beat.
  int valid = __popc(okmask);
  sort_network(v);
  if (paired) {
    valid += __shfl_xor_sync(AEW_FULL, valid, 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const double o1 = __shfl_xor_sync(AEW_FULL, v[31 - i], 1);   // pairs with my v[i]
      const double o2 = __shfl_xor_sync(AEW_FULL, v[i], 1);        // pairs with my v[31 - i]
      const bool t1 = half ? (o1 > v[i]) : (o1 < v[i]);
      const bool t2 = half ? (o2 > v[31 - i]) : (o2 < v[31 - i]);
      v[i] = t1 ? o1 : v[i];
      v[31 - i] = t2 ? o2 : v[31 - i];
    }
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if ((i & stride) == 0) cex(v[i], v[i + stride]);
    }
  }
  const int h = valid >> 1;
  const int cap = paired ? 63 : 31;
  const int rhi = h < cap ? h : cap;
  const int rlo = (valid & 1) ? h : (h > 0 ? h - 1 : 0);
  double lo = 0.0, hi = 0.0;
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    lo = (b == (rlo & 31)) ? v[b] : lo;
    hi = (b == (rhi & 31)) ? v[b] : hi;
  }
  if (paired) {
    const double plo = __shfl_xor_sync(AEW_FULL, lo, 1), phi = __shfl_xor_sync(AEW_FULL, hi, 1);
    lo = ((rlo >> 5) == half) ? lo : plo;
    hi = ((rhi >> 5) == half) ? hi : phi;
  }
  if (valid == 0) return __longlong_as_double(0x7ff8000000000000ll);
  if (z.zs) {
    lo = div_exact(__dsub_rn(lo, z.mean), z.dv);
    hi = div_exact(__dsub_rn(hi, z.mean), z.dv);
  }
  return __dmul_rn(__dadd_rn(lo, hi), 0.5);
}

// Any number of beats, by one thread, by counting: the element of rank r is the one with <= r smaller values and
// > r values that are smaller or equal.  O(n^2) but a dozen instructions of code; only reached beyond 64 beats
// per window, which the 180 bpm screen of the reference rules out for its own window length.
__device__ __noinline__ double column_median_count(const BeatSrc& z, const int* __restrict__ peaks, int n,
                                                   int c, int L, int pre, int post) {
  int valid = 0;
  for (int b = 0; b < n; ++b) { int t = 0; (void)beat_value(z, peaks[b], c, L, pre, post, t); valid += t; }
  if (valid == 0) return __longlong_as_double(0x7ff8000000000000ll);
  const int h = valid >> 1;
  const int rhi = h, rlo = (valid & 1) ? h : h - 1;
  double lo = 0.0, hi = 0.0;
  for (int b = 0; b < n; ++b) {
    int t = 0;
    const double vb = beat_value(z, peaks[b], c, L, pre, post, t);
    if (!t) continue;
    int less = 0, leq = 0;
    for (int j = 0; j < n; ++j) {
      int u = 0;
      const double vj = beat_value(z, peaks[j], c, L, pre, post, u);
      less += (u && vj < vb) ? 1 : 0;
      leq += (u && vj <= vb) ? 1 : 0;
    }
    if (less <= rlo && rlo < leq) lo = vb;
    if (less <= rhi && rhi < leq) hi = vb;
  }
  return __dmul_rn(__dadd_rn(lo, hi), 0.5);
}

// all pre+post columns of one window; every thread of the block calls it
__device__ __forceinline__ void window_median(const BeatSrc& z, const int* __restrict__ peaks, int n,
                                              int L, int pre, int post, double* __restrict__ med,
                                              float* __restrict__ waves) {
  const int S = pre + post;
  if (n > 64) {
    for (int c = threadIdx.x; c < S; c += blockDim.x) {
      const double m = column_median_count(z, peaks, n, c, L, pre, post);
      if (med) med[c] = m;
      if (waves) waves[c] = (float)m;
    }
    return;
  }
  const bool paired = n > 32;
  const int per_pass = paired ? (int)(blockDim.x >> 1) : (int)blockDim.x;
  for (int c0 = 0; c0 < S; c0 += per_pass) {             // warp-uniform trip count
    const int c = c0 + (paired ? (int)(threadIdx.x >> 1) : (int)threadIdx.x);
    const int half = paired ? (int)(threadIdx.x & 1) : 0;
    const bool active = c < S;
    if (!__any_sync(AEW_FULL, active)) continue;
    const double m = column_median_64(z, peaks, n, active ? c : 0, half, active, paired, L, pre, post);
    if (active && half == 0) {
      if (med) med[c] = m;
      if (waves) waves[c] = (float)m;
    }
  }
}

// ---------------------------------------------------------------------- the window kernel
// Most prominent strict local maximum of one QRS run d[0..n) (scipy.signal.find_peaks + peak_prominences
// semantics), by one warp.  Returns the index inside the run, or 0x7fffffff if the run has no peak.
__device__ __noinline__ int run_peak(const double* __restrict__ d, int n, int lane) {
  double best = -1.0;
  int best_i = 0x7fffffff;
  if (n <= 128) {
    // warp-cooperative evaluation (runs of at most 128 samples): lane l holds d[32 m + l]
    double v[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) v[m] = (32 * m + lane < n) ? d[32 * m + lane] : 0.0;
    unsigned long long kv[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) kv[m] = dkey(v[m]);
    // strict local maxima; a plateau (two equal neighbours after a rise) takes the general detector below
    unsigned pkbits[4];
    int plateau = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int i = 32 * m + lane;
      const bool inner = i >= 1 && i < n - 1;
      const double lft = inner ? d[i - 1] : 0.0, rgt = inner ? d[i + 1] : 0.0;
      pkbits[m] = __ballot_sync(AEW_FULL, inner && lft < v[m] && rgt < v[m]);
      plateau |= (inner && lft < v[m] && rgt == v[m]) ? 1 : 0;
    }
    if (__any_sync(AEW_FULL, plateau)) {
      unsigned peakmask[4];
#pragma unroll 1
      for (int m = 0; m < 4; ++m) {
        const int i = 32 * m + lane;
        const double vm = (i < n) ? d[i] : 0.0;
        bool rise = false;
        int pk = i;
        if (i >= 1 && i < n - 1 && d[i - 1] < vm) {          // start of a plateau
          int ahead = i + 1;
          while (ahead < n - 1 && d[ahead] == vm) ++ahead;
          if (d[ahead] < vm) { rise = true; pk = (i + ahead - 1) / 2; }
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
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        bool is_pk = (peakmask[m] & 1u) != 0;
#pragma unroll
        for (int o = 0; o < 4; ++o) is_pk |= (peakmask[o] & (2u << m)) != 0;
        pkbits[m] = __ballot_sync(AEW_FULL, is_pk);
      }
    }
    // Shortcuts before any prominence is formed.  No candidate: no peak.  One candidate: it is the peak.  Several:
    // the HIGHEST candidate g has no higher sample anywhere in the run, so its prominence is h_g - max(min of the
    // run left of it, min right of it); every other candidate's prominence is at most (second-highest candidate)
    // - (minimum of the run), in exact arithmetic and -- rounding being monotonic -- as computed.  If the first
    // exceeds that bound, g is the answer (an R peak between its Q and S dips against the ripples around it: the
    // common case); otherwise, or if the highest value is shared, the candidates are walked one by one below.
    {
      const int cnt = __popc(pkbits[0]) + __popc(pkbits[1]) + __popc(pkbits[2]) + __popc(pkbits[3]);
      if (cnt == 0) return 0x7fffffff;
      if (cnt == 1) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (pkbits[m]) return 32 * m + __ffs((int)pkbits[m]) - 1;
      }
      unsigned long long ck = 0ull;                          // highest candidate key of this lane
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (((pkbits[m] >> lane) & 1u) && kv[m] > ck) ck = kv[m];
      const unsigned long long Mk = ~warp_min_key_raw(~ck);  // maximum over the warp
      int g = -1, ties = 0;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const unsigned b = __ballot_sync(AEW_FULL, ((pkbits[m] >> lane) & 1u) && kv[m] == Mk);
        ties += __popc(b);
        if (b && g < 0) g = 32 * m + __ffs((int)b) - 1;
      }
      if (ties == 1) {
        unsigned long long c2 = 0ull, kl = ~0ull, kr = ~0ull;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int i = 32 * m + lane;
          if (((pkbits[m] >> lane) & 1u) && i != g && kv[m] > c2) c2 = kv[m];
          if (i < n && i <= g && kv[m] < kl) kl = kv[m];
          if (i < n && i >= g && kv[m] < kr) kr = kv[m];
        }
        const double h2 = dkey_inv(~warp_min_key_raw(~c2));
        const double lmin = dkey_inv(warp_min_key_raw(kl)), rmin = dkey_inv(warp_min_key_raw(kr));
        const double hg = dkey_inv(Mk);
        const double prom_g = __dsub_rn(hg, lmin > rmin ? lmin : rmin);
        const double bound = __dsub_rn(h2, lmin < rmin ? lmin : rmin);
        if (prom_g > bound) return g;
      }
    }
    // candidates in ascending order; ONE copy of the prominence code
#pragma unroll 1
    for (int m = 0; m < 4; ++m) {
      unsigned s = m == 0 ? pkbits[0] : (m == 1 ? pkbits[1] : (m == 2 ? pkbits[2] : pkbits[3]));
#pragma unroll 1
      for (; s; s &= s - 1) {
        const int pk = 32 * m + __ffs(s) - 1;
        const double h = d[pk];
        // nearest strictly higher sample on each side (walks of scipy's _peak_prominences stop there)
        int nl = -1, nr = n;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (32 * c < n) {                                  // (uniform) chunks beyond the run hold nothing
            const int i = 32 * c + lane;
            const unsigned hi = __ballot_sync(AEW_FULL, i < n && v[c] > h);
            if (hi) {
              const unsigned below = (32 * c + 31 < pk) ? hi : ((32 * c > pk) ? 0u : (hi & ((1u << (pk - 32 * c)) - 1u)));
              const unsigned above = (32 * c > pk) ? hi : ((32 * c + 31 < pk) ? 0u : (hi & ~((2u << (pk - 32 * c)) - 1u)));
              if (below) { const int q = 32 * c + 31 - __clz((int)below); nl = q > nl ? q : nl; }
              if (above) { const int q = 32 * c + __ffs((int)above) - 1; nr = q < nr ? q : nr; }
            }
          }
        }
        // minima of (nl, pk] and [pk, nr) as order-preserving integer keys: two 32-bit warp reductions each
        unsigned long long kl = ~0ull, kr = ~0ull;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = 32 * c + lane;
          if (i > nl && i <= pk && kv[c] < kl) kl = kv[c];
          if (i >= pk && i < nr && kv[c] < kr) kr = kv[c];
        }
        const double lmin = warp_min_key(kl), rmin = warp_min_key(kr);
        const double prom = __dsub_rn(h, lmin > rmin ? lmin : rmin);
        if (prom > best) { best = prom; best_i = pk; }       // candidates come in ascending order: first max wins
      }
    }
    return best_i;                                           // every lane holds the same best
  }
  for (int i = 1 + lane; i < n - 1; i += 32) {               // long runs (noise): one candidate per lane, walks
    if (d[i - 1] < d[i]) {
      int ahead = i + 1;
      while (ahead < n - 1 && d[ahead] == d[i]) ++ahead;
      if (d[ahead] < d[i]) {
        const int pk = (i + ahead - 1) / 2;                  // plateau midpoint
        const double h = d[pk];
        double lmin = h, rmin = h;
        for (int j = pk; j >= 0 && d[j] <= h; --j) lmin = d[j] < lmin ? d[j] : lmin;
        for (int j = pk; j <= n - 1 && d[j] <= h; ++j) rmin = d[j] < rmin ? d[j] : rmin;
        const double prom = __dsub_rn(h, lmin > rmin ? lmin : rmin);
        if (prom > best || (prom == best && pk < best_i)) { best = prom; best_i = pk; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(AEW_FULL, best, o);
    const int oi = __shfl_xor_sync(AEW_FULL, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  return best_i;
}


// d[k] = hi[k] - lo[k] for the K3_ITEMS consecutive entries behind two positions of the window buffer, read with
// 16-byte loads whatever the parity of the two positions (a thread's strip starts at an even position, so the
// parities are uniform over the block): at a stride of 22 doubles, 8-byte accesses of a warp collide two to a bank
// pair and took twice the shared-memory wavefronts.  f(k, d) consumes the differences in order.
template <bool LO_ODD, bool HI_ODD, class F>
__device__ __forceinline__ void tap_diff(const double* lo, const double* hi, F&& f) {
  const double2* ql = reinterpret_cast<const double2*>(lo - (LO_ODD ? 1 : 0));
  const double2* qh = reinterpret_cast<const double2*>(hi - (HI_ODD ? 1 : 0));
  double2 tl, th;
  tl.x = tl.y = th.x = th.y = 0.0;
  if (LO_ODD) tl = ql[0];
  if (HI_ODD) th = qh[0];
#pragma unroll
  for (int m = 0; m < K3_ITEMS / 2; ++m) {
    double l0, l1, h0, h1;
    if (!LO_ODD) { tl = ql[m]; l0 = tl.x; l1 = tl.y; } else { l0 = tl.y; tl = ql[m + 1]; l1 = tl.x; }
    if (!HI_ODD) { th = qh[m]; h0 = th.x; h1 = th.y; } else { h0 = th.y; th = qh[m + 1]; h1 = th.x; }
    f(2 * m, h0 - l0);
    f(2 * m + 1, h1 - l1);
  }
}
template <class F>
__device__ __forceinline__ void tap_diff_any(const double* W, int lo, int hi, F&& f) {
  if ((lo & 1) == 0) {
    if ((hi & 1) == 0) tap_diff<false, false>(W + lo, W + hi, f); else tap_diff<false, true>(W + lo, W + hi, f);
  } else {
    if ((hi & 1) == 0) tap_diff<true, false>(W + lo, W + hi, f); else tap_diff<true, true>(W + lo, W + hi, f);
  }
}

template <bool FAST>
__global__ void __launch_bounds__(K3_NT, 2) k3_windows(K3Params p, K3Ws ws, K3Leaves lv, unsigned long long* dbg) {
  AEW_DYN_SMEM(smem_raw);
  K3Smem sm;
  k3_smem_layout(smem_raw, &sm);
  const int tid = threadIdx.x;
  const int L = p.L;
  const long long w = blockIdx.x;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
#if !defined(AEW_EMUL) && defined(AEW_K3_TICKS)
  long long t_last = dbg ? clock64() : 0;
#endif
  if (p.win_start[w] < 0) {                             // padding slot of an over-allocated window table
    if (tid == 0) {
      p.keep[w] = 0; p.has_nan[w] = 2; p.n_peaks[w] = 0; ws.route[w] = 0;
      for (int k = 0; k < 4; ++k) p.stats[4 * w + k] = qnan;
    }
    for (int i = tid; i < K3_MAXP; i += K3_NT) p.peaks[w * K3_MAXP + i] = -1;
    return;
  }
  const double* src = p.y + p.win_start[w];
#ifndef AEW_EMUL
  {
    // Blocks are dispatched in index order, two per SM: the window of the block one wave ahead is needed about one
    // block lifetime from now.  Pull it into L2 so that its load phase does not wait on HBM.
    const long long wn = w + K3_PREFETCH_AHEAD;
    if (wn < p.n_win) {
      const long long sn = p.win_start[wn];
      if (sn >= 0) {
        const char* nxt = (const char*)(p.y + sn);
#pragma unroll 1
        for (int off = tid * 128; off < L * 8; off += K3_NT * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + off));
      }
    }
  }
#endif
  const int ks = p.smooth_k, ka = p.avg_k;
  const int ls = ks / 2, la = ka / 2;                  // uniform_filter1d: window [i - k/2, i - k/2 + k)
  const int H = k3_halo(ks, ka);
  double* W = sm.W;
  double* A = W + H;                                    // sample j of the window sits at position j + H of W
  for (int i = tid; i < L / 32 + 2; i += K3_NT) sm.Qw[i] = 0u;

  // ---- load.  ONE bulk copy global -> shared (TMA, completion on an mbarrier) issued by thread 0 where the window
  //      starts on a 16-byte boundary; 8-byte asynchronous copies by all threads otherwise (20 requests per thread
  //      in flight at once, no registers staged)
  const bool tma_in = (L & 1) == 0 && ((unsigned long long)src & 15ull) == 0;
  if (tma_in) {
    if (tid == 0) { mbar_init(sm.mbar, 1); bulk_load(A, src, (unsigned)L * 8u, sm.mbar); }
  } else {
#ifdef AEW_EMUL
    for (int i = tid; i < L; i += K3_NT) A[i] = src[i];
#else
    const unsigned a0 = (unsigned)__cvta_generic_to_shared(A);
#pragma unroll 8
    for (int i = tid; i < L; i += K3_NT)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(a0 + 8u * (unsigned)i), "l"(src + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
  }
  __syncthreads();                                      // window (or its mbarrier), leaf table
  if (tma_in && !mbar_wait(sm.mbar, 0) && tid == 0) atomicOr(p.status, AEW_DEV_LOOKBACK_TIMEOUT);
  // one leaf per group of four threads
  const int depth = lv.depth;
  const int nslots = 1 << depth;
  const int grp = tid >> 2, lj = tid & 3;
  const int slot = grp < nslots ? (int)lv.slot[grp] : grp;
  const int offa = slot < nslots ? (int)lv.off[slot] : 0, sza = slot < nslots ? (int)lv.size[slot] : 0;
  const bool all_leaves = nslots == K3_NT / 4 && lv.all != 0;
  double* zw = p.zwin ? p.zwin + w * (long long)L : nullptr;
  const bool z16 = (L & 1) == 0 && ((unsigned long long)p.zwin & 15ull) == 0;
  int buf = 0;
#define K3_SWEEP(MODE, M, SD, SUM, FLAG)                                                                   \
  {                                                                                                        \
    const LeafOut la_ = leaf_pass(MODE, A, zw, z16, M, SD, offa, sza, lj);                                 \
    SUM = block_combine(la_.r, sza, slot, lj, depth, sm.pw_vals + 128 * buf, sm.pw_valid + 128 * buf, all_leaves); \
    FLAG = la_.flag;                                                                                       \
    buf ^= 1;                                                                                              \
  }

  // fast_stats: no NumPy-order sweeps at all.  The window's mean and sd come out of the first prefix pass below
  // (shifted sums, any order: ~1e-15 from NumPy's), nothing is z-scored here -- the indicator is scale-invariant,
  // the beat kernel takes the R peaks from the unscaled samples, the median kernel z-scores the middle values of its
  // columns -- and nothing is parked.  For filter output that is itself ~1e-9 from SciPy's (the tiled filter).
  constexpr bool fast = FAST;                           // (p.fast_stats, as a template parameter)
  // ---- NaN screen (remove_samples.py:31-33) + sum of the raw window in one sweep
  double sum_raw = 0.0;
  int flag = 0;
  int bad = 0;
  if (!fast) {
    K3_SWEEP(K3_PASS_PLAIN, 0.0, 1.0, sum_raw, flag)
    bad = __syncthreads_or(flag);
  }
  K3_TICK(0)   // load + NaN screen + first sum
  double sum_z = sum_raw;                               // windows that arrive z-scored: the same numbers
  double mean = 0.0, sd0 = 1.0;
  bool parked = false;
  if (!fast && !bad && p.zscore_windows) {
    // ---- per-window z-score (data_extract_vitaldb.py:93), NumPy summation order.  The z-scores go STRAIGHT to
    // the parked copy in global memory ('ecg_sample': the beat, median and delineation kernels read them there);
    // the window in shared memory stays as it came: the QRS indicator below is a comparison of two box sums of
    // |gradient|, invariant under the positive scale 1 / sd, and is taken from the unscaled samples.
    mean = k3_div(sum_raw, (double)L);
    double ssq;
    K3_SWEEP(K3_PASS_DEV2, mean, 1.0, ssq, flag)
    sd0 = k3_sqrt(k3_div(ssq, (double)L));
    K3_SWEEP(K3_PASS_ZSCORE, mean, sd0, sum_z, flag)
    bad = __syncthreads_or(flag);
    parked = true;
    if (tid == 0 && p.zparams) { p.zparams[2 * w] = mean; p.zparams[2 * w + 1] = sd0; }
  }
  // ecg_sd = np.std(window) (remove_samples.py:62) of the z-scored window is 1 up to a few ulp: it decides
  // sd / n_peaks > threshold only when n_peaks * threshold == 1.  With lazy_sd the beat kernel forms
  // it for those windows alone (from the parked copy); without, it is formed here for every window.
  double sd_full = qnan;
  if (!fast && !bad && !p.lazy_sd) {
    const double mz = k3_div(sum_z, (double)L);
    const LeafOut la_ = p.zscore_windows ? leaf_pass_checked(K3_PASS_ZDEV2, A, nullptr, mean, sd0, mz, offa, sza, lj)
                                         : leaf_pass(K3_PASS_DEV2, A, nullptr, false, mz, 1.0, offa, sza, lj);
    const double ssq2 = block_combine(la_.r, sza, slot, lj, depth, sm.pw_vals + 128 * buf, sm.pw_valid + 128 * buf, all_leaves);
    buf ^= 1;
    sd_full = k3_sqrt(k3_div(ssq2, (double)L));
  }
#undef K3_SWEEP
  K3_TICK(1)   // window z-score (+ sum of the z-scored window)
  if (zw && !parked && !fast) {                         // windows the caller z-scored, NaN windows: parked as they came
#pragma unroll 4
    for (int i = tid; i < L; i += K3_NT) zw[i] = A[i];
  }
  if (bad) {
    if (tid == 0) {
      p.keep[w] = 0; p.has_nan[w] = 1; p.n_peaks[w] = 0; ws.route[w] = 0;
      for (int k = 0; k < 4; ++k) p.stats[4 * w + k] = qnan;
    }
    for (int i = tid; i < K3_MAXP; i += K3_NT) p.peaks[w * K3_MAXP + i] = -1;
    return;
  }
  if (tid == 0 && !fast) {                              // what the beat kernel needs besides the bits
    ws.sumz[w] = sum_z;
    ws.sd[w] = sd_full;
    ws.mean[w] = mean;
    ws.sd0[w] = sd0;
    p.has_nan[w] = 0;
    sm.scal[0] = mean; sm.scal[1] = sd0;                // (for the sequential evaluation's restore, if it comes to that)
  }

  // ---- QRS indicator: smoothgrad > 1.5 * avggrad, smoothgrad / avggrad = box filters (scipy "nearest" edges) of
  // |gradient| / smoothgrad.  Evaluated with prefix sums over the WHOLE window IN PLACE: position e of W stands
  // for sample e - H clamped into the window (edge value repeated), so that a box sum is P[hi] - P[lo] everywhere.
  // Thread t owns the strip e = 22 t .. 22 t + 21 in all three passes:
  //   1  W <- P1 = prefix sums of g' = |x[j+1] - x[j-1]| (each thread reads its strip and one sample either side
  //      BEFORE anyone writes; the factor 1/2 of np.gradient's central differences and the two 1/k of the box
  //      filters are exact or common scalings of both sides of the comparison and are folded into ONE constant),
  //   2  s'(e) = P1[e' + ks] - P1[e'] held in REGISTERS over the block scan whose barrier ends all reads of P1,
  //      W <- P2 = prefix sums of s',
  //   3  s' > (1.5 / ka) (P2[e - la - 1 + ka] - P2[e - la - 1]), packed into the bit array.
  // The bits go to the beat kernel (k3_beats), which reads the z-scored samples it needs from the parked copy.
  // Round 1 / early round 2 evaluated this in four segments with their own prefix buffer: 8 block scans and four
  // times the per-thread fixed cost per window, 48 % of the kernel's instructions (profiles/r2_k3_windows.txt).
  int hazard = 0;
  const bool prefix_passes = p.mode != 1;
  if (prefix_passes) {
    const int e0 = tid * K3_ITEMS;
    const int j0 = e0 - H;                              // sample the strip's first position stands for
    const double g0 = 2.0 * fabs(__dsub_rn(A[1], A[0])), gL = 2.0 * fabs(__dsub_rn(A[L - 1], A[L - 2]));
    const double left = e0 > 0 ? W[e0 - 1] : 0.0;
    const double right = e0 + K3_ITEMS < K3_WCAP ? W[e0 + K3_ITEMS] : 0.0;
    if (fast) {
      // sums of d = x - x[0] and d^2 over the strip's samples (shifted: no cancellation however far the window sits
      // from zero); a NaN or an infinity anywhere in the window ends up in the totals
      const double x0 = A[0];
      double s1a = 0.0, s1b = 0.0, s2a = 0.0, s2b = 0.0;
      if (__all_sync(AEW_FULL, j0 >= 0 && j0 + K3_ITEMS <= L)) {
        const double2* Wq = reinterpret_cast<const double2*>(W + e0);
#pragma unroll
        for (int m = 0; m < K3_ITEMS / 2; ++m) {
          const double2 v = Wq[m];
          const double dx = v.x - x0, dy = v.y - x0;
          s1a += dx; s2a += dx * dx;
          s1b += dy; s2b += dy * dy;
        }
      } else {
#pragma unroll 2
        for (int k = 0; k < K3_ITEMS; ++k) {
          const int j = j0 + k;
          const double d = W[e0 + k] - x0;              // (positions outside the window hold anything)
          const bool in = j >= 0 && j < L;
          s1a += in ? d : 0.0;
          s2a += in ? d * d : 0.0;
        }
      }
      double s1 = s1a + s1b, s2 = s2a + s2b;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(AEW_FULL, s1, o);
        s2 += __shfl_xor_sync(AEW_FULL, s2, o);
      }
      if ((tid & 31) == 0) { sm.pw_vals[tid >> 5] = s1; sm.pw_vals[K3_NW + (tid >> 5)] = s2; }
    }
    __syncthreads();                                    // every strip's neighbours are in registers
    if (fast && tid < 32) {
      // warp 0 finishes the window's statistics while the others are in their strips: lanes 0..15 hold the warps'
      // sums of d, lanes 16..31 those of d^2
      double t = tid < 2 * K3_NW ? sm.pw_vals[tid] : 0.0;
#pragma unroll
      for (int o = K3_NW / 2; o > 0; o >>= 1) t += __shfl_xor_sync(AEW_FULL, t, o);
      const double S2 = __shfl_sync(AEW_FULL, t, 16);
      if (tid == 0) {
        const double dL = (double)L;
        const double md = k3_div(t, dL);
        const double var = k3_div(S2 - t * md, dL);     // (sum d^2 - (sum d)^2 / L) / L
        const double sdw = k3_sqrt(var);
        const bool ok = var > 0.0 && sdw > 0.0 && sdw < 1e300;   // (a NaN fails every comparison)
        mean = A[0] + md;
        sm.scal[0] = mean; sm.scal[1] = sdw; sm.scal[5] = ok ? 0.0 : 1.0;
        ws.sumz[w] = qnan; ws.sd[w] = qnan; ws.mean[w] = mean; ws.sd0[w] = sdw;
        if (p.zparams) { p.zparams[2 * w] = mean; p.zparams[2 * w + 1] = sdw; }
        if (ok) p.has_nan[w] = 0;
      }
    }
    double2* Wp = reinterpret_cast<double2*>(W + e0);
    double run = 0.0;
    double loc[K3_ITEMS];                               // the strip's local prefix sums: held over the block scan, so
                                                        // that the strip is written ONCE, with its base added
    if (__all_sync(AEW_FULL, j0 >= 1 && j0 + K3_ITEMS - 1 <= L - 2)) {
      double prv = left;
      double2 cur = Wp[0];
#pragma unroll
      for (int m = 0; m < K3_ITEMS / 2; ++m) {
        double2 nxt;
        if (m + 1 < K3_ITEMS / 2) nxt = Wp[m + 1];
        else { nxt.x = right; nxt.y = 0.0; }
        run += fabs(cur.y - prv);   loc[2 * m] = run;       // position e0 + 2m: samples either side
        run += fabs(nxt.x - cur.x); loc[2 * m + 1] = run;   // position e0 + 2m + 1
        prv = cur.y; cur = nxt;
      }
    } else {                                            // strips that touch an edge of the window (warp-uniform)
      double prv = left, cur = W[e0];
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) {
        const double nxt = k + 1 < K3_ITEMS ? W[e0 + k + 1] : right;
        const int j = j0 + k;
        double g = fabs(nxt - prv);
        g = j <= 0 ? g0 : g;                            // np.gradient is one-sided at the two ends (x 2 here)
        g = j >= L - 1 ? gL : g;
        run += g;
        loc[k] = run;
        prv = cur; cur = nxt;
      }
    }
    double base = block_excl_scan(run, sm.wtot);        // (its barrier follows every strip's reads of the samples)
    if (tid == K3_NT - 1) sm.scal[4] = base + run;      // sum of g' over all positions: scale of the tie tolerance
#pragma unroll
    for (int m = 0; m < K3_ITEMS / 2; ++m) {
      double2 o;
      o.x = base + loc[2 * m]; o.y = base + loc[2 * m + 1];
      Wp[m] = o;
    }
    __syncthreads();                                    // P1 complete
    if (fast && sm.scal[5] != 0.0) {                    // a NaN in the window, or a constant one (no z-score): removed
      if (tid == 0) {
        p.keep[w] = 0; p.has_nan[w] = 1; p.n_peaks[w] = 0; ws.route[w] = 0;
        for (int k = 0; k < 4; ++k) p.stats[4 * w + k] = qnan;
      }
      for (int i = tid; i < K3_MAXP; i += K3_NT) p.peaks[w * K3_MAXP + i] = -1;
      return;
    }
    K3_TICK(3)   // |gradient| prefix sums
    double sv[K3_ITEMS];
    run = 0.0;
    if (__all_sync(AEW_FULL, j0 >= 0 && j0 + K3_ITEMS - 1 <= L - 1)) {
      tap_diff_any(W, e0 - ls - 1, e0 - ls - 1 + ks, [&](int k, double d) { sv[k] = d; run += d; });
    } else {
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) {
        int i = j0 + k;
        i = i < 0 ? 0 : (i > L - 1 ? L - 1 : i);        // smoothgrad's own edge value repeated
        const double* q = W + (i - ls - 1 + H);
        sv[k] = q[ks] - q[0];
        run += sv[k];
      }
    }
    base = block_excl_scan(run, sm.wtot + K3_NW);       // (its barrier ends all reads of P1)
    {
      double r = base;
#pragma unroll
      for (int m = 0; m < K3_ITEMS / 2; ++m) {
        double2 o;
        r += sv[2 * m];     o.x = r;
        r += sv[2 * m + 1]; o.y = r;
        Wp[m] = o;
      }
    }
    __syncthreads();                                    // P2 complete
    K3_TICK(4)   // smoothgrad prefix sums
    // outputs: samples 0 <= j < L; avggrad box of position e = P2[e - la - 1 + ka] - P2[e - la - 1]
    const double c15 = k3_div(1.5, (double)ka);
    const double tol_abs = 2e-12 * sm.scal[4];
    int klo = -j0, khi = L - j0;
    klo = klo < 0 ? 0 : (klo > K3_ITEMS ? K3_ITEMS : klo);
    khi = khi < klo ? klo : (khi > K3_ITEMS ? K3_ITEMS : khi);
    unsigned qbits = 0;
    // smoothgrad of the strip comes back as differences of its own prefix sums (an error of one ulp of the prefix,
    // a hundredth of the tie tolerance): the 22 values of pass 2 need not stay in registers through this pass
    double prevP = e0 > 0 ? W[e0 - 1] : 0.0;
    double2 own;
    own.x = own.y = 0.0;
    if (__all_sync(AEW_FULL, klo == 0 && khi == K3_ITEMS)) {
      tap_diff_any(W, e0 - la - 1, e0 - la - 1 + ka, [&](int k, double d) {
        if ((k & 1) == 0) own = Wp[k >> 1];
        const double cur = (k & 1) ? own.y : own.x;
        const double sk = cur - prevP;
        prevP = cur;
        const double thr = d * c15;
        qbits |= (unsigned)(sk > thr) << k;
        hazard |= (fabs(sk - thr) <= 1e-9 * (sk + thr) + tol_abs) ? 1 : 0;
      });
    } else if (khi > klo) {
      const double* pl = W + e0 - la - 1;
#pragma unroll
      for (int k = 0; k < K3_ITEMS; ++k) {
        const double cur = W[e0 + k];
        const double sk = cur - prevP;
        prevP = cur;
        if (k >= klo && k < khi) {
          const double thr = (pl[k + ka] - pl[k]) * c15;
          qbits |= (unsigned)(sk > thr) << k;
          hazard |= (fabs(sk - thr) <= 1e-9 * (sk + thr) + tol_abs) ? 1 : 0;
        }
      }
    }
    if (khi > klo) {
      // the strip's bits (22 at most) land in one or two words of the bit array
      const int s0 = j0 + klo;                          // first sample of the strip that is an output
      const unsigned bits = qbits >> klo;
      if (bits) {
        const int wq = s0 >> 5, sh = s0 & 31;
        atomicOr((int*)sm.Qw + wq, (int)(bits << sh));
        if (sh > 32 - K3_ITEMS && (bits >> (32 - sh))) atomicOr((int*)sm.Qw + wq + 1, (int)(bits >> (32 - sh)));
      }
    }
    hazard = __syncthreads_or(hazard);                  // (also ends all reads of P2)
    K3_TICK(5)   // indicator + hazard vote
    if (p.mode == 2) hazard = 0;
    if (p.mode == 3) hazard = 1;                        // test hook: every window takes the fallback transition
    if (hazard && tid == 0 && p.hazard_count) atomicAdd(p.hazard_count, 1ull);

  }
  if (p.mode == 1 || hazard) {
    if (prefix_passes) {
      // ---- the z-scored samples come back for the sequential evaluation (rare): the window is read again and the
      //      z-score sweep repeated -- the same operations on the same operands, the same bits
#ifdef AEW_EMUL
      for (int i = tid; i < L; i += K3_NT) A[i] = src[i];
#else
      const unsigned a0 = (unsigned)__cvta_generic_to_shared(A);
#pragma unroll 4
      for (int i = tid; i < L; i += K3_NT)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(a0 + 8u * (unsigned)i), "l"(src + i) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
      __syncthreads();
      if (p.zscore_windows) {
        (void)leaf_pass_checked(K3_PASS_ZSCORE, A, A, sm.scal[0], sm.scal[1], 0.0, offa, sza, lj);
        __syncthreads();
      }
    } else if (p.zscore_windows) {
      // mode 1: the window in shared memory is still unscaled; the sequential evaluation wants SciPy's operands
      __syncthreads();                                  // (thread 0's sm.scal)
      (void)leaf_pass_checked(K3_PASS_ZSCORE, A, A, sm.scal[0], sm.scal[1], 0.0, offa, sza, lj);
      __syncthreads();
    }
    if (tid == 0) {
      const double dks = (double)ks, dka = (double)ka;
      double* ring = sm.ring;
      double t = 0.0;
      for (int e = 0; e < ks; ++e) {
        int j = e - ls; j = j < 0 ? 0 : (j > L - 1 ? L - 1 : j);
        t = __dadd_rn(t, absgrad(A, j, L));
      }
      int js = 0;                                        // smoothgrad is known up to sample js
      ring[0] = __ddiv_rn(t, dks);
      auto smooth_at = [&](int j) -> double {            // advances the first filter as far as needed
        while (js < j) {
          ++js;
          int jin = js + ks - 1 - ls, jout = js - 1 - ls;
          jin = jin > L - 1 ? L - 1 : jin;
          jout = jout < 0 ? 0 : jout;
          t = __dadd_rn(t, __dsub_rn(absgrad(A, jin, L), absgrad(A, jout, L)));
          ring[js & (K3_RING - 1)] = __ddiv_rn(t, dks);
        }
        return ring[j & (K3_RING - 1)];
      };
      double t2 = 0.0;
      for (int e = 0; e < ka; ++e) {
        int j = e - la; j = j < 0 ? 0 : (j > L - 1 ? L - 1 : j);
        t2 = __dadd_rn(t2, smooth_at(j));
      }
      unsigned acc = smooth_at(0) > __dmul_rn(1.5, __ddiv_rn(t2, dka)) ? 1u : 0u;
      for (int i = 1; i < L; ++i) {
        int jin = i + ka - 1 - la, jout = i - 1 - la;
        jin = jin > L - 1 ? L - 1 : jin;
        jout = jout < 0 ? 0 : jout;
        const double sin_ = smooth_at(jin);              // (advance first: jout and i are then inside the ring)
        t2 = __dadd_rn(t2, __dsub_rn(sin_, smooth_at(jout)));
        const unsigned q = smooth_at(i) > __dmul_rn(1.5, __ddiv_rn(t2, dka)) ? 1u : 0u;
        if ((i & 31) == 0) { sm.Qw[(i >> 5) - 1] = acc; acc = 0u; }
        acc |= q << (i & 31);
      }
      sm.Qw[(L - 1) >> 5] = acc;
    }
    __syncthreads();
  }
  K3_TICK(6)   // sequential fallback (hazard windows / mode 1)

  // ---- the indicator leaves for the beat kernel (k3_beats): one bit per sample, and what np.std will need there
  {
    unsigned* qout = ws.qbits + (size_t)blockIdx.x * K3_QWS;
    for (int i = tid; i < (L + 31) / 32 + 1; i += K3_NT) qout[i] = sm.Qw[i];
  }
  K3_TICK(7)   // bits out
}

// ---- the beat kernel: QRS runs -> R peaks -> HR / RMSSD / np.std -> keep decision, from the indicator bits of the
// window kernel and the parked z-scored samples.  A kernel of its own with SMALL blocks (128 threads, ~17 KB of
// shared memory, eight blocks per SM): this part of the path is a chain of short, mostly one-warp steps on a few
// dozen numbers.  Inside the window kernel it ran while the block's 90 KB window buffer sat idle, two windows per
// SM, and its ~2,500 cold instructions were fetched from L2 once per window (profiles/README.md, "instruction
// fetch"); here eight windows per SM overlap each other's latencies and the code stays in the instruction cache.
#define K3P_NT 128
#define K3P_NW (K3P_NT / 32)
#define K3P_WPT ((K3_QWS + K3P_NT - 1) / K3P_NT)      // indicator words per thread (3)

__device__ __forceinline__ int beats_excl_scan_int(int v, int* wtot) {   // 128 threads, one barrier
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(AEW_FULL, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) wtot[warp] = incl;
  __syncthreads();
  int before = 0;
#pragma unroll
  for (int k = 0; k < K3P_NW; ++k) before += (k < warp) ? wtot[k] : 0;
  return before + (incl - v);
}

struct SumDev2 {                                        // (z[i] - m)^2 of samples in global memory
  const double* z; double m;
  __device__ __forceinline__ double operator()(long long i, bool) { const double d = __dsub_rn(z[i], m); return __dmul_rn(d, d); }
};

struct SumPlainG { const double* x; __device__ __forceinline__ double operator()(long long i, bool) { return x[i]; } };
struct SumZ {                                           // z = (x[i] - mean) / sd, or (z - m2)^2
  const double* x; double mean; ExactDiv dv; double m2; int dev2;
  __device__ __forceinline__ double operator()(long long i, bool) {
    const double zz = div_exact(__dsub_rn(x[i], mean), dv);
    if (!dev2) return zz;
    const double d = __dsub_rn(zz, m2);
    return __dmul_rn(d, d);
  }
};

__global__ void __launch_bounds__(K3P_NT, 8) k3_beats(K3Params p, K3Ws ws) {
  __shared__ unsigned s_q[K3_QWS + 2];
  __shared__ unsigned short s_beg[K3_MAXRUN], s_end[K3_MAXRUN];
  __shared__ double s_run[K3P_NW][136];               // a run's samples, one row per warp
  __shared__ double s_vals[128], s_dsq[K3_MAXP], s_scal[4];
  __shared__ unsigned char s_valid[128];
  __shared__ int s_ipeaks[K3_MAXP], s_idiff[K3_MAXP], s_isort[K3_MAXP], s_wtot[K3P_NW + 1], s_iscal[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = p.L;
  const long long w = blockIdx.x;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  if (p.has_nan[w] != 0) return;                        // NaN / padding windows: the window kernel has written their rows
  // fast_stats: the unscaled samples.  Strict local maxima and the ranking of prominences are those of the z-scored
  // samples (a positive affine map) unless two prominences tie to the last bit.
  const double* z = (p.zscore_windows && !p.fast_stats) ? p.zwin + w * (long long)L : p.y + p.win_start[w];
  const int nwords = (L + 31) >> 5;
  {
    const unsigned* qin = ws.qbits + (size_t)w * K3_QWS;
    for (int i = tid; i < nwords + 1; i += K3P_NT) s_q[i] = i < nwords ? qin[i] : 0u;
  }
  __syncthreads();

  // ---- QRS runs: beg where ~q[i] & q[i+1], end where q[i] & ~q[i+1] (ordered compaction: each thread takes
  //      K3P_WPT consecutive words of the bit array, one block scan of the packed counts)
  int nbeg, nend;
  {
    unsigned bb[K3P_WPT], eb[K3P_WPT];
    int mine = 0;
#pragma unroll
    for (int u = 0; u < K3P_WPT; ++u) {
      const int wi = tid * K3P_WPT + u;
      unsigned qb = 0, nx = 0;
      if (wi < nwords) { qb = s_q[wi]; nx = s_q[wi + 1] & 1u; }
      const unsigned nextq = (qb >> 1) | (nx << 31);      // flag of sample i + 1
      int nval = L - 1 - 32 * wi;                         // transitions exist for 0 <= i < L - 1
      nval = nval < 0 ? 0 : (nval > 32 ? 32 : nval);
      const unsigned vmask = nval >= 32 ? 0xffffffffu : ((1u << nval) - 1u);
      bb[u] = ~qb & nextq & vmask; eb[u] = qb & ~nextq & vmask;
      mine += __popc(bb[u]) | (__popc(eb[u]) << 16);
    }
    const int packed = beats_excl_scan_int(mine, s_wtot);
    if (tid == K3P_NT - 1) s_iscal[0] = packed + mine;    // totals
    int ob = packed & 0xffff, oe = packed >> 16;
#pragma unroll
    for (int u = 0; u < K3P_WPT; ++u) {
      const int wi = tid * K3P_WPT + u;
      unsigned b = bb[u], e = eb[u];
      while (b) {
        const int k = __ffs(b) - 1; b &= b - 1;
        if (ob < K3_MAXRUN) s_beg[ob] = (unsigned short)(32 * wi + k);
        ++ob;
      }
      while (e) {
        const int k = __ffs(e) - 1; e &= e - 1;
        if (oe < K3_MAXRUN) s_end[oe] = (unsigned short)(32 * wi + k);
        ++oe;
      }
    }
    __syncthreads();
    nbeg = s_iscal[0] & 0xffff;
    nend = s_iscal[0] >> 16;
    if (nbeg > K3_MAXRUN || nend > K3_MAXRUN) {          // more runs than the tables hold: reported, not guessed
      if (tid == 0) atomicOr(p.status, AEW_DEV_RUN_OVERFLOW);
      nbeg = nbeg > K3_MAXRUN ? K3_MAXRUN : nbeg;
      nend = nend > K3_MAXRUN ? K3_MAXRUN : nend;
    }
  }
  // ends that precede the first beg are dropped (at most one, runs alternate)
  const int skip = (nbeg > 0 && nend > 0 && s_end[0] <= s_beg[0]) ? 1 : 0;
  int num = nend - skip < nbeg ? nend - skip : nbeg;
  if (nbeg == 0) num = 0;
  const unsigned short* endp = s_end + skip;

  // ---- min_len = mean(end - beg) * 0.4 (integer sum is exact in any order); every warp computes it for itself
  double min_len = 0.0;
  {
    unsigned part = 0;                                  // <= num * L < 2^31
    for (int i = lane; i < num; i += 32) part += (unsigned)(endp[i] - s_beg[i]);
    const unsigned tot = __reduce_add_sync(AEW_FULL, part);
    if (num > 0) min_len = __dmul_rn(k3_div((double)tot, (double)num), 0.4);
  }
  __syncthreads();                                      // beg[] is overwritten by the candidates below

  // ---- per run: most prominent strict local maximum (scipy.signal.find_peaks semantics); cand[r] replaces beg[r].
  //      A run of up to 128 samples is copied into the warp's row first (one coalesced read of the window).
  //      (Staging ALL runs of the window at once with asynchronous copies was measured: 1 ms per 179,000 windows
  //      SLOWER -- 31 KB of shared memory leave seven blocks per SM instead of eight, and the extra scan and copy
  //      loop cost more than the overlapped round trips save.)
  for (int r = warp; r < num; r += K3P_NW) {
    const int b = s_beg[r], e = endp[r];
    const int n = e - b;
    int best_i = 0x7fffffff;
    if (!((double)n < min_len)) {
      if (n <= 128) {
        double* row = s_run[warp];
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (32 * m + lane < n) row[32 * m + lane] = z[b + 32 * m + lane];
        __syncwarp();
        best_i = run_peak(row, n, lane);
      } else {
        best_i = run_peak(z + b, n, lane);                // long runs (noise): straight from global memory
      }
    }
    __syncwarp();
    if (lane == 0) s_beg[r] = best_i == 0x7fffffff ? (unsigned short)K3_NONE : (unsigned short)(b + best_i);
  }
  __syncthreads();
  const unsigned short* cand = s_beg;

  // ---- everything between the candidates and the keep decision is a few hundred operations on <= 72 numbers:
  //      warp 0 does it alone, ordered by warp barriers
  if (warp == 0) {
    int np = 0;
    bool chained = false;
    if (num <= 32) {
      // refractory chain, all candidates at once: if every candidate is more than the refractory period behind
      // the previous one, every one is kept (the sequential rule compares with the last KEPT one, which then is
      // the previous one)
      const int c = lane < num ? (int)cand[lane] : (int)K3_NONE;
      const bool have = c != (int)K3_NONE;
      const unsigned mask = __ballot_sync(AEW_FULL, have);
      const unsigned below = mask & ((1u << lane) - 1u);
      const int prev = __shfl_sync(AEW_FULL, c, below ? 31 - __clz((int)below) : 0);
      const bool ok = !have || c - (below ? prev : 0) > p.refractory;
      if (__all_sync(AEW_FULL, ok)) {
        np = __popc(mask);
        if (have) s_ipeaks[__popc(below)] = c;
        chained = true;
      }
    }
    if (!chained) {                                        // sequential by definition; all lanes walk it alike
      int last = 0;
      for (int r = 0; r < num; ++r) {
        const int c = cand[r];
        if (c != (int)K3_NONE && c - last > p.refractory) {
          if (np < K3_MAXP && lane == 0) s_ipeaks[np] = c;
          ++np;
          last = c;
        }
      }
    }
    if (np > K3_MAXP) { if (lane == 0) atomicOr(p.status, AEW_DEV_RUN_OVERFLOW); np = K3_MAXP; }
    __syncwarp();
    // remove_samples.py:53-54 keeps 0 < p < L (always true here: p > refractory)
    // HR = 60 fs / median(RR) and RMSSD (remove_samples.py:85-121)
    const int nd = np - 1;
    for (int t = lane; t < nd; t += 32) {
      const int d0 = s_ipeaks[t + 1] - s_ipeaks[t];
      s_idiff[t] = d0;
      if (t + 1 < nd) {
        const double cur = k3_div((double)(s_ipeaks[t + 2] - s_ipeaks[t + 1]), p.rr_div);
        const double prv = k3_div((double)d0, p.rr_div);
        const double dd = __dsub_rn(cur, prv);
        s_dsq[t] = __dmul_rn(dd, dd);
      }
    }
    __syncwarp();
    for (int t = lane; t < nd; t += 32) {                  // rank sort of <= 71 integers
      const int mine = s_idiff[t];
      int rank = 0;
      for (int j = 0; j < nd; ++j) {
        const int o = s_idiff[j];
        rank += (o < mine) || (o == mine && j < t);
      }
      s_isort[rank] = mine;
    }
    SumPlain fq{s_dsq};                                    // NumPy leaf order over np-2 <= 70 values
    const double ssq = pw_leaf_sum8(fq, 0, np >= 3 ? np - 2 : 0, lane & 7);
    __syncwarp();
    if (lane == 0) {
      double hr = qnan, rms = qnan;
      if (np >= 2) {
        const double med = (nd & 1) ? (double)s_isort[nd / 2]
                                    : __dmul_rn(__dadd_rn((double)s_isort[nd / 2 - 1], (double)s_isort[nd / 2]), 0.5);
        hr = k3_div(p.hr_num, med);
      }
      if (np >= 3) rms = k3_sqrt(k3_div(ssq, (double)(np - 2)));
      s_iscal[2] = np;
      s_scal[0] = hr;
      s_scal[1] = rms;
    }
  }
  __syncthreads();
  const int np = s_iscal[2];
  if (tid < K3_MAXP) p.peaks[w * K3_MAXP + tid] = tid < np ? s_ipeaks[tid] : -1;

  // ---- ecg_sd = np.std(window) (remove_samples.py:62) and the keep decision (:70-74).  The window is z-scored
  // (by the window kernel, or by the caller), so sd = 1 up to a few ulp and sd / n_peaks > threshold is decided by
  // n_peaks alone unless n_peaks * threshold == 1 (20 beats at 0.05: SURVEY.md C.1).  lazy_sd: the NumPy-order
  // sweep over the window runs only for those windows (here, over the parked samples); the others report sd = NaN
  // and take the decision that every sd within 1e-6 of 1 gives.  (A window the CALLER z-scored is trusted to be
  // z-scored only with lazy_sd set.)
  const double npt = __dmul_rn((double)np, p.sd_threshold);
  const bool need_sd = !p.lazy_sd || np == 0 || fabs(npt - 1.0) < 1e-6;
  double sd = qnan;
  if (need_sd) {
    if (!p.lazy_sd) {
      sd = ws.sd[w];
    } else if (p.fast_stats) {
      // Here the LAST BIT of np.std decides, so this window (one in fifteen) gets the whole NumPy-order chain after
      // all: its mean and sd as np.mean / np.std form them, the z-scores from those on the fly, their sum, and the
      // sum of their squared deviations -- the value, and the decision, of the chain without fast_stats.
      const int dep = pw_depth(L);
      SumPlainG fx{z};
      const double mean_e = k3_div(pw_block_sum(fx, L, dep, s_vals, s_valid), (double)L);
      SumDev2 fd{z, mean_e};
      const double sd_e = k3_sqrt(k3_div(pw_block_sum(fd, L, dep, s_vals, s_valid), (double)L));
      SumZ fz{z, mean_e, make_exact_div(sd_e), 0.0, 0};
      fz.m2 = k3_div(pw_block_sum(fz, L, dep, s_vals, s_valid), (double)L);
      fz.dev2 = 1;
      const double ssq2 = pw_block_sum(fz, L, pw_depth(L), s_vals, s_valid);
      sd = k3_sqrt(k3_div(ssq2, (double)L));
    } else {
      SumDev2 f{z, k3_div(ws.sumz[w], (double)L)};
      const double ssq2 = pw_block_sum(f, L, pw_depth(L), s_vals, s_valid);
      sd = k3_sqrt(k3_div(ssq2, (double)L));
    }
  }
  if (tid == 0) {
    const double hr = s_scal[0], rms = s_scal[1];
    const double sdn = need_sd ? k3_div(sd, (double)np) : qnan;
    const bool sd_rule = need_sd ? (sdn > p.sd_threshold) : (npt < 1.0);      // 1 / np > threshold
    const bool remove = (hr < p.hr_lo) || (hr > p.hr_hi) || isnan(hr) || (rms > p.pvc_threshold) || sd_rule;
    p.keep[w] = remove ? 0 : 1;
    p.n_peaks[w] = np;
    // the median kernel of this window: the key-sorting ones take 3..32 / 33..64 beats whose only incomplete sections
    // can be the first and the last beat's; each median kernel reads this one byte and leaves if it is not its own
    int route = 0;
    if (!remove) {
      const bool plain = np >= 3 && np <= 64 && s_ipeaks[1] >= p.pre && s_ipeaks[np - 2] + p.post <= L;
      route = plain ? (np <= 32 ? 1 : 2) : 3;
    }
    ws.route[w] = (unsigned char)route;
    p.stats[4 * w + 0] = hr; p.stats[4 * w + 1] = rms; p.stats[4 * w + 2] = sd; p.stats[4 * w + 3] = sdn;
  }
}

// ---- K3b: median beat of kept windows (median_sample.py:6-54) -> np_waves row.
// A kernel of its own: the 32-input sorting network is 25 KB of straight-line code; inside the screen kernel it
// was fetched from L2 once per window (see the note on instruction fetch above), here it stays in the
// instruction cache while the blocks stream through the windows.  One thread per beat column; the samples are
// re-read from global memory (L2 for the most part: the screen kernel has just read them) and z-scored on the
// fly with the window's mean / sd saved by the screen kernel.
// Blocks of 160 threads take half of a window's 320 columns each: three of them fit an SM (registers), and their
// start-up round trips to global memory overlap.
#define K3B_NT 160

// Windows of 3..32 beats whose only incomplete sections can be the first and the last beat's (all but a handful): the
// column median from 32-BIT KEYS.  The high word of a double, made order-preserving, with its low five bits
// replaced by the slot number, sorts like the double down to 2^-15 relative; the 32 keys of a column are sorted with
// 191 x 2 integer min / max instead of 191 x (DSETP + 4 SEL) on register pairs, and only the four values around the
// middle come back as doubles (re-read: they are in L1) to be ordered exactly.  The wanted ranks are FIXED: the
// slots without a sample are padded with -inf and +inf in equal numbers (to within one), so that the middle of the
// padded column -- sorted positions 15 and 16 -- is the middle of its samples.  Exactness: position 14..17 hold, in
// key order, the samples of true rank 14..17 unless three or more samples around the middle agree in their key's 27
// value bits; then (once in ~10^4 columns) the window is handed to k3_median.
#define K3K_NEG 0x000fffc0u                             // key of -inf / +inf with (six) slot bits cleared; every finite
#define K3K_POS 0xfff00000u                             // double lies strictly between them
__device__ __forceinline__ unsigned hikey(unsigned hi) {
  return hi ^ ((unsigned)((int)hi >> 31) | 0x80000000u);
}
// NS = 32 slots (3..32 beats, five slot bits) or 64 (33..64 beats, six: keys down to 2^-14 relative)
template <int NS>
__global__ void __launch_bounds__(K3B_NT, NS == 32 ? 6 : 4) k3_median_keys(K3Params p, K3Ws ws, int nparts) {
  constexpr unsigned SLOT = NS - 1;                       // mask of the slot bits
  constexpr int H = NS / 2;                               // the column's middle: sorted positions H - 1 and H
  __shared__ int s_peaks[NS];
  const int S = p.pre + p.post;
  const long long w = blockIdx.x / nparts;
  if (ws.route[w] != (NS == 32 ? 1 : 2)) return;          // (the other instance's, k3_median's, or nobody's)
  const int c = (int)(blockIdx.x - w * nparts) * K3B_NT + (int)threadIdx.x;
  const int n = p.n_peaks[w];
  const long long start = p.win_start[w];
  const bool derive = p.zscore_windows && (!p.zwin || p.fast_stats);
  const double mean = derive ? ws.mean[w] : 0.0;
  const double sd0 = derive ? ws.sd0[w] : 1.0;
  if (threadIdx.x < NS) s_peaks[threadIdx.x] = p.peaks[w * K3_MAXP + threadIdx.x];
  __syncthreads();
  const int L = p.L, pre = p.pre, post = p.post;
  const double* zb = (derive || !p.zwin) ? p.y + start : p.zwin + w * (long long)L;
  const unsigned* zh = reinterpret_cast<const unsigned*>(zb) + 1;     // high words (little-endian)
  const bool active = c < S;
  const int cc = active ? c : 0;
  // slot 0 <- the first beat, slot NS-1 <- the last one, slots 1 .. n-2 <- the beats between (whole sections),
  // slots n-1 .. NS-2 <- padding
  const int pk0 = s_peaks[0], pkl = s_peaks[n - 1];
  const int s00 = pk0 - pre < 0 ? 0 : pk0 - pre, e00 = pk0 + post > L ? L : pk0 + post;
  const int s0l = pkl - pre, e0l = pkl + post > L ? L : pkl + post;
  const bool ok0 = s00 + cc < e00, okl = s0l + cc < e0l;
  // padding alternates -inf, +inf from slot n-1 on; the first one's sign is chosen so that the column ends up with
  // as many -inf as +inf, give or take one
  const int P = NS - n;
  const bool neg_first = !(!ok0 && okl);
  const int a_pad = (neg_first ? (P + 1) >> 1 : P >> 1) + (ok0 ? 0 : 1);    // -inf in the column
  const int b_pad = (neg_first ? P >> 1 : (P + 1) >> 1) + (okl ? 0 : 1);    // +inf
  unsigned k[NS];
  k[0] = ok0 ? zh[2 * (s00 + cc)] : 0u;
  k[NS - 1] = okl ? zh[2 * (s0l + cc)] : 0u;
#pragma unroll
  for (int b = 1; b < NS - 1; ++b) k[b] = b <= n - 2 ? zh[2 * (s_peaks[b] - pre + cc)] : 0u;
  k[0] = ok0 ? (hikey(k[0]) & ~SLOT) : K3K_NEG;
  k[NS - 1] = (okl ? (hikey(k[NS - 1]) & ~SLOT) : K3K_POS) | SLOT;
  const unsigned pad_even = neg_first ? K3K_NEG : K3K_POS, pad_odd = neg_first ? K3K_POS : K3K_NEG;
#pragma unroll
  for (int b = 1; b < NS - 1; ++b) {
    const unsigned real = hikey(k[b]) & ~SLOT;
    const unsigned pad = ((b - (n - 1)) & 1) ? pad_odd : pad_even;
    k[b] = (b <= n - 2 ? real : pad) | (unsigned)b;
  }
  sort_network(k);
  // positions H-2 .. H+1 as doubles, in exact order
  double v4[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const unsigned kq = k[H - 2 + q];
    const unsigned tk = kq & ~SLOT;
    const int slot = (int)(kq & SLOT);
    const int beat = slot == NS - 1 ? n - 1 : slot;
    const int pk = s_peaks[beat];
    const int s0 = pk - pre < 0 ? 0 : pk - pre;
    double x = __longlong_as_double(tk == K3K_NEG ? (long long)0xfff0000000000000ull : 0x7ff0000000000000ll);
    if (tk != K3K_NEG && tk != K3K_POS) x = zb[s0 + cc];
    v4[q] = x;
  }
  cex(v4[0], v4[1]); cex(v4[2], v4[3]); cex(v4[0], v4[2]); cex(v4[1], v4[3]); cex(v4[1], v4[2]);
  // three samples or more with one key around the middle: their order among the positions is not known
  const unsigned ta = k[H - 3] & ~SLOT, tb = k[H - 1] & ~SLOT, tc = k[H] & ~SLOT, td = k[H + 2] & ~SLOT;
  const bool unsure = (ta == tb && tb != K3K_NEG && tb != K3K_POS) || (tc == td && tc != K3K_NEG && tc != K3K_POS);
  if (__any_sync(AEW_FULL, active && unsure)) {
    if ((threadIdx.x & 31) == 0) ws.route[w] = 3;          // (both halves of the window are redone by k3_median)
  }
  double lo = a_pad > b_pad ? v4[2] : v4[1];
  double hi = b_pad > a_pad ? v4[1] : v4[2];
  if (derive) {
    const ExactDiv dv = make_exact_div(sd0);
    lo = div_exact(__dsub_rn(lo, mean), dv);
    hi = div_exact(__dsub_rn(hi, mean), dv);
  }
  const double m = __dmul_rn(__dadd_rn(lo, hi), 0.5);
  if (active) {
    if (p.median) p.median[w * S + c] = m;
    if (p.waves) p.waves[w * S + c] = (float)m;
  }
}

// the general form: any number of beats, any incomplete sections; the windows k3_median_keys does not take
__global__ void __launch_bounds__(K3B_NT, 3) k3_median(K3Params p, K3Ws ws, int nparts, int keys_done) {
  __shared__ int s_peaks[K3_MAXP];
  const int S = p.pre + p.post;                          // section length int(0.64 fs): nparts = ceil(S / K3B_NT)
  const long long w = blockIdx.x / nparts;
  const int c_lo = (int)(blockIdx.x - w * nparts) * K3B_NT;
  {
    const int route = ws.route[w];
    if (route == 0 || (keys_done && route != 3)) return;  // removed, or done by a keys kernel
  }
  // everything the block needs to get going, in one round trip
  const int keep = p.keep[w];
  const int n = p.n_peaks[w];
  const long long start = p.win_start[w];
  const bool derive = p.zscore_windows && (!p.zwin || p.fast_stats);
  const double mean = derive ? ws.mean[w] : 0.0;
  const double sd0 = derive ? ws.sd0[w] : 1.0;
  if (threadIdx.x < K3_MAXP) s_peaks[threadIdx.x] = p.peaks[w * K3_MAXP + threadIdx.x];
  if (!keep) return;                                     // rows of removed windows stay undefined, as before
  __syncthreads();
  BeatSrc src;                                           // the z-scored window if the screen kernel parked it
  src.z = derive || !p.zwin ? p.y + start : p.zwin + w * (long long)p.L;
  src.zs = derive;
  src.mean = mean;
  src.dv = make_exact_div(sd0);
  double* med = p.median ? p.median + w * S : nullptr;
  float* waves = p.waves ? p.waves + w * S : nullptr;
  if (n > 64) {
    const int c = c_lo + threadIdx.x;
    if (c >= S) return;
    const double m = column_median_count(src, s_peaks, n, c, p.L, p.pre, p.post);
    if (med) med[c] = m;
    if (waves) waves[c] = (float)m;
    return;
  }
  const bool paired = n > 32;
  const int per_pass = paired ? K3B_NT / 2 : K3B_NT;
#pragma unroll 1
  for (int c0 = 0; c0 < K3B_NT; c0 += per_pass) {
    const int c = c_lo + c0 + (paired ? (int)(threadIdx.x >> 1) : (int)threadIdx.x);
    const int half = paired ? (int)(threadIdx.x & 1) : 0;
    const bool active = c < S;                           // (the last part of a section that is no multiple of 160)
    if (!__any_sync(AEW_FULL, active)) continue;
    const double m = column_median_64(src, s_peaks, n, active ? c : 0, half, active, paired, p.L, p.pre, p.post);
    if (active && half == 0) {
      if (med) med[c] = m;
      if (waves) waves[c] = (float)m;
    }
  }
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
  BeatSrc src;
  src.z = z; src.zs = 0; src.mean = 0.0; src.dv = make_exact_div(1.0);
  window_median(src, pk, n, L, pre, post, median + w * S, nullptr);
}

}  // namespace

static unsigned long long* g_k3_dbg = nullptr;   // device counters, allocated once when AEW_K3_TIMING is set
static unsigned long long* k3_timing_buffer() {
#ifndef AEW_EMUL
  if (!g_k3_dbg && getenv("AEW_K3_TIMING")) {
    cudaMalloc(&g_k3_dbg, 16 * sizeof(unsigned long long));
    cudaMemset(g_k3_dbg, 0, 16 * sizeof(unsigned long long));
  }
#endif
  return g_k3_dbg;
}

extern "C" int aew_windows_max_peaks(void) { return K3_MAXP; }
extern "C" int aew_windows_max_len(void) { return K3_MAXL; }
extern "C" size_t aew_windows_workspace_bytes(long long n_win) {
  return (size_t)(n_win > 0 ? n_win : 0) * (4 * sizeof(double) + K3_QWS * sizeof(unsigned) + 1) + 64;
}

extern "C" int aew_windows_beats_f64(const K3Params* hp, void* stream) {
  if (!hp) return AEW_ERR_ARG;
  K3Params p = *hp;
  if (!p.y || !p.win_start || !p.keep || !p.has_nan || !p.n_peaks || !p.peaks || !p.stats || !p.status)
    return AEW_ERR_ARG;
  if (p.n_win == 0) return AEW_OK;
  if (p.L < 16 || p.L > K3_MAXL || p.n_win < 0 || p.n_win > 0x7fffffffll) return AEW_ERR_UNSUPPORTED;
  if (p.smooth_k < 1 || p.avg_k < 1) return AEW_ERR_UNSUPPORTED;
  // the window with the two box filters' halo must fit the window buffer, the sequential evaluation's ring must span
  // the second box filter
  if (k3_ext_len(p.L, p.smooth_k, p.avg_k) > K3_WCAP || p.avg_k + 2 > K3_RING) return AEW_ERR_UNSUPPORTED;
  if (pw_depth(p.L) > 7 || p.pre < 0 || p.post < 1 || p.pre + p.post > K3_MAXS || p.smooth_k < 1 || p.avg_k < 1 || p.smooth_k > p.L ||
      p.avg_k > p.L || p.refractory < 1 || (p.L / (p.refractory + 1) + 1) > K3_MAXP || p.n_win > 0x3fffffffll)
    return AEW_ERR_UNSUPPORTED;
  const size_t smem = k3_smem_layout(nullptr, nullptr);
  if (smem > 113 * 1024) return AEW_ERR_UNSUPPORTED;    // two blocks per SM
  // the beat kernel reads the z-scored samples of the QRS runs from the parked copy: with zscore_windows it is required
  if (p.zscore_windows && !p.zwin && !p.fast_stats) return AEW_ERR_ARG;
  // fast_stats: per-window z-score without NumPy's summation order; nothing is parked, np.std only where it decides
  if (p.fast_stats && (!p.zscore_windows || !p.lazy_sd || p.mode == 1 || p.zwin)) return AEW_ERR_ARG;
  if (!p.workspace || p.workspace_bytes < aew_windows_workspace_bytes(p.n_win)) return AEW_ERR_WORKSPACE;
  K3Ws ws;
  {
    unsigned char* base = (unsigned char*)p.workspace;
    base += (16 - ((size_t)base & 15)) & 15;
    ws.sumz = (double*)base;
    ws.sd = ws.sumz + p.n_win;
    ws.mean = ws.sd + p.n_win;
    ws.sd0 = ws.mean + p.n_win;
    ws.qbits = (unsigned*)(ws.sd0 + p.n_win);
    ws.route = (unsigned char*)(ws.qbits + (size_t)p.n_win * K3_QWS);
  }
  // (per call: the attribute is per DEVICE, and the call costs a fraction of a microsecond)
  if (AEW_SET_MAX_SMEM(k3_windows<false>, 113 * 1024) != 0) return (int)cudaGetLastError();
  if (AEW_SET_MAX_SMEM(k3_windows<true>, 113 * 1024) != 0) return (int)cudaGetLastError();
  K3Leaves lv;
  {
    lv.depth = pw_depth(p.L);
    const int ns = 1 << lv.depth;
    lv.all = 1;
    int cls[128], n0 = 0, n1 = 0, list0[128], list1[128];
    for (int i = 0; i < 128; ++i) {
      long long off = 0, size = 0;
      if (i < ns) pw_node(p.L, lv.depth, (unsigned)i, off, size);
      lv.off[i] = (unsigned short)off; lv.size[i] = (unsigned char)size; lv.slot[i] = (unsigned char)i;
      if (i < ns && size == 0) lv.all = 0;
      cls[i] = (int)((off >> 3) & 1);
      if (i < ns) { if (cls[i]) list1[n1++] = i; else list0[n0++] = i; }
    }
    if (ns == 128 && n0 == n1)                          // groups 2q, 2q + 1 <- the q-th leaf of either residue
      for (int q = 0; q < 64; ++q) { lv.slot[2 * q] = (unsigned char)list0[q]; lv.slot[2 * q + 1] = (unsigned char)list1[q]; }
  }
  if (p.fast_stats) {
    AEW_LAUNCH(k3_windows<true>, (unsigned)p.n_win, K3_NT, smem, (cudaStream_t)stream, p, ws, lv, k3_timing_buffer());
  } else {
    AEW_LAUNCH(k3_windows<false>, (unsigned)p.n_win, K3_NT, smem, (cudaStream_t)stream, p, ws, lv, k3_timing_buffer());
  }
  AEW_LAUNCH(k3_beats, (unsigned)p.n_win, K3P_NT, 0, (cudaStream_t)stream, p, ws);
  if (p.median || p.waves) {
    const int nparts = (p.pre + p.post + K3B_NT - 1) / K3B_NT;
    const int use_keys = getenv("AEW_K3_MEDIAN_GENERAL") ? 0 : 1;   // (measurement switch: the general kernel alone)
    if (use_keys) {
      AEW_LAUNCH(k3_median_keys<32>, (unsigned)(nparts * p.n_win), K3B_NT, 0, (cudaStream_t)stream, p, ws, nparts);
      AEW_LAUNCH(k3_median_keys<64>, (unsigned)(nparts * p.n_win), K3B_NT, 0, (cudaStream_t)stream, p, ws, nparts);
    }
    AEW_LAUNCH(k3_median, (unsigned)(nparts * p.n_win), K3B_NT, 0, (cudaStream_t)stream, p, ws, nparts, use_keys);
  }
  return (int)cudaGetLastError();
}

// debugging aid (AEW_K3_TIMING=1): cycles thread 0 of every block spent per phase of k3_windows, summed over
// blocks; copies and clears the 16 counters (synchronises the device)
extern "C" int aew_windows_timing_dump(unsigned long long* h_out16) {
#ifndef AEW_EMUL
  if (!g_k3_dbg || !h_out16) return AEW_ERR_ARG;
  cudaDeviceSynchronize();
  cudaMemcpy(h_out16, g_k3_dbg, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaMemset(g_k3_dbg, 0, 16 * sizeof(unsigned long long));
  return AEW_OK;
#else
  (void)h_out16;
  return AEW_ERR_UNSUPPORTED;
#endif
}

// windows: [n_win][L] contiguous, peaks: CSR (peak_off[n_win+1]); at most 128 peaks per window.
extern "C" int aew_median_beats_f64(const double* windows, long long n_win, int L, const int* peaks,
                                    const long long* d_peak_off, int max_peaks_per_window, int pre, int post,
                                    double* median, double* aligned, void* stream) {
  if (!windows || !peaks || !d_peak_off || !median || n_win < 0 || L < 1) return AEW_ERR_ARG;
  if (max_peaks_per_window > 128 || pre < 0 || post < 1 || pre + post > K3_MAXS || n_win > 0x7fffffffll)
    return AEW_ERR_UNSUPPORTED;
  if (n_win == 0) return AEW_OK;
  AEW_LAUNCH(k4_median, (unsigned)n_win, K3_NT, 0, (cudaStream_t)stream, windows, L, peaks, d_peak_off, pre, post,
             median, aligned);
  return (int)cudaGetLastError();
}
