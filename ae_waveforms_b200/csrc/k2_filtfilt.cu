// k2_filtfilt.cu -- zero-phase IIR filtering of NaN-delimited segments (K2).
//
// Replaces data_load_preprocess/filters.py:15-55 (`butterworth`): the output is
// NaN where the input is NaN; every maximal non-NaN segment longer than
// padlen = 3*max(len(a),len(b)) gets scipy.signal.filtfilt(b, a, seg)
// (odd extension by padlen, zi-scaled initial state, forward DF2T pass,
// reversal, second pass, reversal, trim -- SURVEY.md appendix B.1); shorter
// segments are copied through (filters.py:50-51).  The optional (mean, sd)
// arguments fuse the whole-record z-score of data_extract_vitaldb.py:64 into the
// load, so the z-scored record is never materialised.
//
// Kernels
//   k2_seg_find   one block per record; finds the segments with a block-wide
//                 "latest start" max-scan, appends (start, length, record,
//                 first tile) to the segment table, fills the tile->segment map
//                 and writes NaN to the output at NaN inputs.  Records whose
//                 non-NaN count equals their length (known from K1) skip the scan.
//   k2_pass<REV>  persistent blocks (K2_BPS = 4 per SM, 128 threads each) pull 1,664-sample
//                 tiles off an atomic counter into a 4-slot ring of shared-memory buffers:
//                 while tile n is computed, tile n+1 arrives by a 1-D bulk copy (TMA,
//                 cp.async.bulk + mbarrier) and tile n-1 leaves by a bulk store, all issued
//                 by one thread.  Thread t runs the recursion over its own 13-sample chunk
//                 from a ZERO state (stride 13 doubles = conflict-free).  The recursion is carried in
//                 MODAL form (two complex first-order sections, k2_tables.cpp) so
//                 that the state transition is diagonal: the chunk end states are
//                 combined by a warp-shuffle Kogge-Stone scan that costs two
//                 complex multiplications per step, warps are chained, and the
//                 state entering the tile comes from a bounded look-back over the
//                 published end states of the previous tiles (|p|^1664 ~ 2e-4 for
//                 the 1 Hz high-pass, so six predecessors suffice, one lane of the
//                 look-back warp each; the depth is computed per filter on the
//                 host).  Outputs are then corrected in
//                 place with y[j] += G[j] . s (zero-input response) and stored
//                 coalesced.
//                 REV=false is the forward pass (x -> tmp), REV=true the backward
//                 pass over the reversed forward output (tmp -> out).
//   k2_exact      one thread per segment, the literal sequential DF2T order
//                 without FMA: bit-identical to SciPy, used by the parity suite
//                 (mode 1) to show the rest of the chain is bit-exact end to end.
//
// Algorithmic HBM bytes: 16 B/sample per pass, 32 B/sample per filtfilt
// (SURVEY.md section 8d).  Roofline: HBM; FP64 work is ~17 FMA/sample/pass.
#include <stdio.h>
#include <stdlib.h>
#include "aew_rt.h"
#include "aew_math.cuh"
#include "k2_layout.h"
#include "aew_bulk.cuh"

namespace {

// ---- workspace header (int64 slots)
enum { H_SEG_COUNT = 0, H_TILE_TOTAL = 1, H_WORK_FWD = 2, H_WORK_BWD = 3, H_STATUS = 4, H_WORK_EXACT = 5, H_SLOTS = 8 };

struct K2Ws {
  long long* hdr;       // [H_SLOTS]
  long long* segs;      // [max_seg][4] : g0, n, rec, tile_base
  double* vext;         // [max_seg][16] forward outputs over the right extension
  long long* tile_desc; // [max_tiles][4] per tile: g0, n, (record << 32) | segment, tile index in the segment
  long long max_seg, max_tiles;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline size_t k2_ws_layout(void* base, long long max_seg, long long max_tiles, K2Ws* w) {
  size_t o = 0;
  unsigned char* p = (unsigned char*)base;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 128); return p ? p + at : (unsigned char*)nullptr; };
  K2Ws t;
  t.hdr = (long long*)take(H_SLOTS * sizeof(long long));
  t.segs = (long long*)take((size_t)max_seg * 4 * sizeof(long long));
  t.vext = (double*)take((size_t)max_seg * 16 * sizeof(double));
  t.tile_desc = (long long*)take((size_t)max_tiles * 4 * sizeof(long long));
  t.max_seg = max_seg;
  t.max_tiles = max_tiles;
  if (w) *w = t;
  return o;
}

// ------------------------------------------------------------ segment discovery
#define SEG_THREADS 256
#define SEG_ITEMS 8

__global__ void __launch_bounds__(SEG_THREADS) k2_seg_find(const double* __restrict__ x,
                                                           const long long* __restrict__ rec_beg, const long long* __restrict__ rec_end,
                                                           const long long* __restrict__ nonnan_count,
                                                           double* __restrict__ nan_fill, K2Ws ws) {
  __shared__ long long s_warp[SEG_THREADS / 32];
  __shared__ long long s_carry;
  const long long r = blockIdx.x;
  const long long base = rec_beg[r], n = rec_end[r] - base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);

  auto append = [&](long long start, long long len) {
    const long long idx = (long long)atomicAdd((unsigned long long*)(ws.hdr + H_SEG_COUNT), 1ull);
    if (idx >= ws.max_seg) { atomicOr((int*)(ws.hdr + H_STATUS), AEW_DEV_SEG_OVERFLOW); return; }
    const long long nt = (len + K2_T - 1) / K2_T;
    const long long tb = (long long)atomicAdd((unsigned long long*)(ws.hdr + H_TILE_TOTAL), (unsigned long long)nt);
    long long* s = ws.segs + 4 * idx;
    s[0] = base + start; s[1] = len; s[2] = r;
    if (tb + nt > ws.max_tiles) {
      atomicOr((int*)(ws.hdr + H_STATUS), AEW_DEV_TILE_OVERFLOW);
      s[3] = -1;
      return;
    }
    s[3] = tb;
    for (long long t = 0; t < nt; ++t) {
      long long* td = ws.tile_desc + 4 * (tb + t);
      td[0] = base + start; td[1] = len; td[2] = (r << 32) | idx; td[3] = t;
    }
  };

  if (n <= 0) return;
  if (nonnan_count && nonnan_count[r] == n) {      // no NaN in this record: one segment, no scan
    if (tid == 0) append(0, n);
    return;
  }
  if (tid == 0) s_carry = -1;
  __syncthreads();
  const double* xr = x + base;
  for (long long blk0 = 0; blk0 < n; blk0 += (long long)SEG_THREADS * SEG_ITEMS) {
    const long long p0 = blk0 + (long long)tid * SEG_ITEMS;
    // local walk: last segment start inside my items (or -1)
    bool prev = (p0 > 0 && p0 - 1 < n) ? !isnan(xr[p0 - 1]) : false;
    bool ok[SEG_ITEMS];
    long long my_last = -1;
#pragma unroll
    for (int k = 0; k < SEG_ITEMS; ++k) {
      const long long p = p0 + k;
      ok[k] = (p < n) ? !isnan(xr[p]) : false;
      if (p < n && !ok[k] && nan_fill) nan_fill[base + p] = qnan;
      const bool pv = k == 0 ? prev : ok[k - 1];
      if (ok[k] && !pv) my_last = p;
    }
    // exclusive max-scan of my_last over the block
    long long incl = my_last;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long o = __shfl_up_sync(AEW_FULL, incl, d);
      if (lane >= d && o > incl) incl = o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    long long before = s_carry;                     // latest start in earlier iterations / warps
    for (int w = 0; w < warp; ++w) if (s_warp[w] > before) before = s_warp[w];
    long long excl = __shfl_up_sync(AEW_FULL, incl, 1);
    if (lane == 0) excl = -1;
    if (before > excl) excl = before;
    // emit the segments that END inside my items
    long long cur = excl;
#pragma unroll
    for (int k = 0; k < SEG_ITEMS; ++k) {
      const long long p = p0 + k;
      if (p >= n) break;
      const bool pv = k == 0 ? prev : ok[k - 1];
      if (ok[k] && !pv) cur = p;
      const bool nx = (p + 1 < n) ? (k + 1 < SEG_ITEMS ? ok[k + 1] : !isnan(xr[p + 1])) : false;
      if (ok[k] && !nx) append(cur, p + 1 - cur);
    }
    __syncthreads();
    if (tid == SEG_THREADS - 1) {
      long long c = s_carry;
      for (int w = 0; w < SEG_THREADS / 32; ++w) if (s_warp[w] > c) c = s_warp[w];
      s_carry = c;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------ modal-form helpers
// s = (u1.re, u1.im, u2.re, u2.im);  u' = p u + x;  y = d x + GC . s   (k2_layout.h)
struct Modal { double d, p1r, p1i, p2r, p2i, g0, g1, g2, g3; };

// one sample (fast path: FMA allowed, results within 1e-6 relative of SciPy's)
__device__ __forceinline__ double modal_step(const Modal& c, double x, double s[4]) {
  const double y = fma(c.g3, s[3], fma(c.g2, s[2], fma(c.g1, s[1], fma(c.g0, s[0], c.d * x))));
  const double n0 = fma(c.p1r, s[0], fma(-c.p1i, s[1], x));
  const double n1 = fma(c.p1r, s[1], c.p1i * s[0]);
  const double n2 = fma(c.p2r, s[2], fma(-c.p2i, s[3], x));
  const double n3 = fma(c.p2r, s[3], c.p2i * s[2]);
  s[0] = n0; s[1] = n1; s[2] = n2; s[3] = n3;
  return y;
}

// acc += P (.) v   (two complex multiplications; P = (p1.re, p1.im, p2.re, p2.im))
__device__ __forceinline__ void cmul_acc(const double* __restrict__ P, const double v[4], double acc[4]) {
  acc[0] = fma(P[0], v[0], fma(-P[1], v[1], acc[0]));
  acc[1] = fma(P[0], v[1], fma(P[1], v[0], acc[1]));
  acc[2] = fma(P[2], v[2], fma(-P[3], v[3], acc[2]));
  acc[3] = fma(P[2], v[3], fma(P[3], v[2], acc[3]));
}

// SciPy's initial DF2T state zi*scale, mapped to modal coordinates
__device__ __forceinline__ void modal_initial(const double* tab, double scale, double s[4]) {
  double z[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) z[i] = tab[K2_TAB_ZI + i] * scale;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double* row = tab + K2_TAB_PZ + 4 * i;
    s[i] = fma(row[3], z[3], fma(row[2], z[2], fma(row[1], z[1], row[0] * z[0])));
  }
}

// z-score on load: (v - mean) / sd with IEEE division (aew_math.cuh), bit-identical to NumPy
__device__ __forceinline__ double zs(double v, bool on, double mean, const ExactDiv& sd) {
  return on ? div_exact(__dsub_rn(v, mean), sd) : v;
}
// the tiled pass multiplies by 1/sd instead (two operations per sample instead of ten): one more rounding of
// relative size 1e-16 in a pass that is within ~1e-9 of SciPy anyway; outputs that are COPIES of z-scored samples
// (segments too short to filter) and the exact mode keep the division
__device__ __forceinline__ double zsf(double v, bool on, double mean, double inv_sd) {
  return on ? (v - mean) * inv_sd : v;
}

// ------------------------------------------------------------ tiled pass
struct PassArgs {
  const double* src;      // forward: x (raw or previous stage); backward: tmp (forward output)
  const double* mean;     // optional per-record z-score (forward only)
  const double* sd;
  double* dst;            // forward: tmp; backward: out
  double* passthrough;    // forward: out (short segments are copied here)
  const double* tab;      // device coefficient table (k2_layout.h)
  K2Ws ws;
  int bulk_ok;            // src and dst base pointers are 16-byte aligned: bulk (TMA) copies allowed
  // Coefficients every thread uses with compile-time indices travel in the kernel parameters: they become
  // constant-bank operands of the FMAs instead of shared-memory loads (52 + 20 broadcast loads per thread and tile)
  Modal cf;               // the filter in modal form
  double cG[4 * K2_C];    // zero-input response rows G[j] (K2_TAB_G)
  double cPK[20];         // p^(C 2^k), k = 0..4 (K2_TAB_PK)
  unsigned long long* dbg; // optional phase-cycle counters (AEW_K2_TIMING=1), else null
  int run_tiles;          // tiles of a segment a block works through in a row (k2_pass)
};

// (bulk asynchronous copies and their mbarrier: aew_bulk.cuh)

struct TileDesc {
  long long g0, n, rec, k, t0, gstart;   // gstart: first global sample index covered by the tile (ascending order)
  int seg, cnt;
  bool bulk;                             // full tile whose global range is 16-byte aligned
};

// raw[4] = one entry of the tile table written by k2_seg_find: g0, n, (rec << 32) | seg, k
template <bool REV>
__device__ __forceinline__ TileDesc describe_tile(const long long raw[4], int bulk_ok, int pad) {
  TileDesc d;
  d.g0 = raw[0]; d.n = raw[1];
  d.rec = raw[2] >> 32;
  d.seg = (int)(raw[2] & 0xffffffffll);
  d.k = raw[3];
  d.t0 = d.k * K2_T;
  const long long remain = d.n - d.t0;
  d.cnt = remain < K2_T ? (int)remain : K2_T;
  d.gstart = REV ? d.g0 + d.n - d.t0 - d.cnt : d.g0 + d.t0;
  d.bulk = bulk_ok && d.n > pad && d.cnt == K2_T && ((d.gstart & 1) == 0);
  return d;
}

#define K2_PRODUCER 32               // thread that claims tiles and issues the bulk copies (warp 1, lane 0)

#ifdef AEW_EMUL
#define K2_TICK(slot)
#define K2_TICK_T(thread, slot, reset)
#else
// cycles a given thread spends between a reset point and here
#define K2_TICK_T(thread, slot, reset)                                  \
  if (a.dbg && tid == (thread)) {                                       \
    const long long now_ = clock64();                                   \
    if (!(reset)) atomicAdd(a.dbg + (slot), (unsigned long long)(now_ - tick_)); \
    tick_ = now_;                                                       \
  }
#define K2_TICK(slot)                                                   \
  if (a.dbg && tid == 0) {                                              \
    const long long now_ = clock64();                                   \
    atomicAdd(a.dbg + (slot), (unsigned long long)(now_ - tick_));      \
    tick_ = now_;                                                       \
  }
#endif

// The order in which a block works through tiles.  A RUN is K2 tiles [m R, (m + 1) R) of one segment (R =
// PassArgs::run_tiles): a block takes a run as a whole and carries the filter state from tile to tile itself, so
// that no tile ever waits for another block.  The state entering a run's first tile (m > 0) comes from a WARM-UP:
// the block first works through the `lookback` tiles in front of the run without storing anything, from a zero
// state -- |p^T|^lookback < 1e-18, the same truncation the look-back over published states used to make.  Runs are
// cut relative to the SEGMENT, so a record's output does not depend on where its tiles sit in the batch's table.
// Blocks claim CHUNKS of R consecutive table entries and own the runs that START inside their chunk (a run's last
// tiles may lie in the next chunk; that chunk's owner skips them).
struct TileSeq {
  long long next, warm_end, run_end, chunk_end;
};

// number of tiles of the segment a table entry belongs to
__device__ __forceinline__ long long tiles_of(long long n) { return (n + K2_T - 1) / K2_T; }

// next tile of the sequence (>= n_tiles: the work is used up); *flags: K2_WARM = warm-up tile (no output),
// K2_FRESH = first tile a block works on for a run (the state entering it is zero unless it is a segment's first)
#define K2_WARM 1
#define K2_FRESH 2
__device__ __forceinline__ long long next_tile(TileSeq& q, const long long* __restrict__ tile_desc, long long* work,
                                               long long n_tiles, long long R, long long D, int* flags) {
  int fresh = 0;
  if (q.next == q.run_end) {
    fresh = K2_FRESH;
    long long g = q.run_end;                           // the entry behind the run just finished
    for (;;) {
      if (g >= q.chunk_end) {                          // chunk used up: claim the next one
        const long long r = (long long)atomicAdd((unsigned long long*)work, 1ull);
        g = r * R;
        if (g >= n_tiles) { q.next = q.run_end = q.chunk_end = n_tiles; *flags = 0; return n_tiles; }
        q.chunk_end = g + R < n_tiles ? g + R : n_tiles;
        const long long k = tile_desc[4 * g + 3], left = tiles_of(tile_desc[4 * g + 1]) - k;
        const long long to_run = (R - k % R) % R;      // entries up to the next run boundary of this segment
        g += to_run < left ? to_run : left;            // (or up to the next segment, whichever comes first)
        continue;                                      // (g may have reached the end of the chunk)
      }
      break;
    }
    const long long k = tile_desc[4 * g + 3], left = tiles_of(tile_desc[4 * g + 1]) - k;
    const long long w = k < D ? k : D;
    q.next = g - w;
    q.warm_end = g;
    q.run_end = g + (R < left ? R : left);
  }
  *flags = fresh | (q.next < q.warm_end ? K2_WARM : 0);
  return q.next++;
}

#define K2_AHEAD 2                    // tiles the producer's copies run ahead of the tile being worked on

template <bool REV>
__global__ void __launch_bounds__(K2_NT, K2_BPS) k2_pass(PassArgs a) {
  long long tick_ = 0;
  (void)tick_;
  AEW_DYN_SMEM(smem_raw);
  double* stage0 = (double*)smem_raw;                 // [K2_STAGES][K2_T] tile ring
  double* s_tab = stage0 + K2_STAGES * K2_T;          // [K2_TAB_LEN]
  double* s_W = s_tab + K2_TAB_LEN;                   // [K2_NW][4] warp totals
  double* s_in = s_W + 4 * K2_NW;                     // [4] state entering a segment's first tile
  double* s_chain = s_in + 4;                         // [2][4] filter state at the end of the tile worked on last
  __shared__ unsigned long long s_bar[K2_STAGES];
  __shared__ long long s_tile[K2_STAGES];
  __shared__ int s_flags[K2_STAGES];                  // K2_WARM | K2_FRESH of the tile in the slot
  // A tile's table entry and its record's z-score parameters travel with the tile: the producer thread, which has the
  // entry in registers an iteration ahead, leaves both next to the ring slot (no thread starts an iteration with two
  // dependent global round trips -- entry, then mean / sd of its record -- in front of the recursion).
  __shared__ long long s_desc[K2_STAGES][4];
  __shared__ double s_zs[K2_STAGES][4];               // mean, sd, 1 / sd of the tile's record (forward pass with z-score)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < K2_TAB_LEN; i += K2_NT) s_tab[i] = a.tab[i];
  if (tid == 0)
    for (int s = 0; s < K2_STAGES; ++s) mbar_init(&s_bar[s], 1);
  __syncthreads();
  const Modal& cf = a.cf;
  const int pad = (int)s_tab[K2_TAB_PAD];
  const long long lookback = (long long)s_tab[K2_TAB_LOOKBACK];
  const long long R = a.run_tiles;
  const bool zon = (!REV) && a.mean != nullptr;
  long long* work = a.ws.hdr + (REV ? H_WORK_BWD : H_WORK_FWD);
  long long n_tiles = a.ws.hdr[H_TILE_TOTAL];
  if (n_tiles > a.ws.max_tiles) n_tiles = a.ws.max_tiles;

  // The producer thread fills ring slot (it + K2_AHEAD) % K2_STAGES while tile `it` is worked on: table entry,
  // z-score parameters, flags, and the bulk copy of the samples.  idn / rawn / flagsn = the tile it will place next
  // (table entry loaded one iteration before it is used).
  TileSeq seq = {0, 0, 0, 0};
  long long idn = 0, rawn[4] = {0, 0, 0, 0};
  int flagsn = 0;
  auto place = [&](int slot, bool wait_store) {        // (producer thread only)
    s_tile[slot] = idn;
    s_flags[slot] = flagsn;
    if (idn < n_tiles) {
      const TileDesc dn = describe_tile<REV>(rawn, a.bulk_ok, pad);
      if (dn.bulk) {
        // the slot's previous occupant was handed to a bulk store K2_STAGES - K2_AHEAD iterations ago; one bulk
        // group is committed per iteration, so "all but the latest group read" covers it
        if (wait_store) bulk_store_wait_read(1);
        bulk_load(stage0 + slot * K2_T, a.src + dn.gstart, K2_T * 8u, &s_bar[slot]);
      }
      for (int i = 0; i < 4; ++i) s_desc[slot][i] = rawn[i];
      const double sdn = zon ? a.sd[dn.rec] : 1.0;
      s_zs[slot][0] = zon ? a.mean[dn.rec] : 0.0; s_zs[slot][1] = sdn; s_zs[slot][2] = zon ? 1.0 / sdn : 1.0;
      idn = next_tile(seq, a.ws.tile_desc, work, n_tiles, R, lookback, &flagsn);
      if (idn < n_tiles)
        for (int i = 0; i < 4; ++i) rawn[i] = a.ws.tile_desc[4 * idn + i];
    }
  };
  if (tid == K2_PRODUCER) {
    idn = next_tile(seq, a.ws.tile_desc, work, n_tiles, R, lookback, &flagsn);
    if (idn < n_tiles)
      for (int i = 0; i < 4; ++i) rawn[i] = a.ws.tile_desc[4 * idn + i];
    for (int s = 0; s < K2_AHEAD; ++s) place(s, false);
  }
  __syncthreads();
  unsigned parity_bits = 0;                          // bit s: phase parity of ring slot s's mbarrier

#ifndef AEW_EMUL
  tick_ = clock64();
#endif
  for (int it = 0;; ++it) {
    const int st = it % K2_STAGES;
    const long long tile = s_tile[st];
    if (tile >= n_tiles) break;                        // (slots are filled in order: nothing is in flight behind it)
    const int flags = s_flags[st];
    K2_TICK(0)

    // =================== recursion over the tile from a zero state, scan of the chunk end states ===================
    long long rw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) rw[i] = s_desc[st][i];
    const TileDesc d = describe_tile<REV>(rw, a.bulk_ok, pad);
    const double mean = s_zs[st][0];
    const double inv_sd = s_zs[st][2];
    double* sm = stage0 + st * K2_T;
    const bool filt = d.n > pad;                       // else filters.py:50-51: too short, copied through
    const int dir = (REV && d.bulk) ? -1 : 1;
    const int base = (REV && d.bulk) ? K2_T - 1 : 0;
    double* mine = sm + base + dir * (tid * K2_C);
    double E[4] = {0.0, 0.0, 0.0, 0.0};
    double y[K2_C];
    if (!filt) {
      if (!REV) {
        const ExactDiv sd = make_exact_div(s_zs[st][1]);
        for (long long i = tid; i < d.n; i += K2_NT) a.passthrough[d.g0 + i] = zs(a.src[d.g0 + i], zon, mean, sd);
      }
    } else {
      if (d.bulk) {
        if (!mbar_wait(&s_bar[st], (parity_bits >> st) & 1u) && tid == 0)   // bytes of the bulk copy have landed
          atomicOr((int*)(a.ws.hdr + H_STATUS), AEW_DEV_LOOKBACK_TIMEOUT);
        parity_bits ^= 1u << st;
      } else {
        if (tid == K2_PRODUCER) bulk_store_wait_read(0);   // the slot may still be read by an earlier bulk store
        __syncthreads();
        for (int j = tid; j < K2_T; j += K2_NT) {      // cooperative copy, zero padded
          double v = 0.0;
          if (j < d.cnt) {
            const long long p = d.t0 + j;
            v = REV ? a.src[d.g0 + d.n - 1 - p] : a.src[d.g0 + p];
          }
          sm[j] = v;
        }
        __syncthreads();
      }
      K2_TICK(6)   // tile data ready (mbarrier wait / cooperative copy)
      // first tile of a segment: run the odd extension (padlen samples) from SciPy's zi-scaled state
      if (d.k == 0 && tid == 0) {
        double z[4];
        if (!REV) {
          const double* xs = a.src + d.g0;
          const double x0 = zsf(xs[0], zon, mean, inv_sd);
          modal_initial(s_tab, 2.0 * x0 - zsf(xs[pad], zon, mean, inv_sd), z);
          for (int e = 0; e < pad; ++e) modal_step(cf, 2.0 * x0 - zsf(xs[pad - e], zon, mean, inv_sd), z);
        } else {
          const double* ve = a.ws.vext + 16ll * d.seg;   // forward outputs over the right extension
          modal_initial(s_tab, ve[pad - 1], z);
          for (int e = 0; e < pad; ++e) modal_step(cf, ve[pad - 1 - e], z);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s_in[i] = z[i];
      }
      // phase 1: zero-state recursion over my chunk; the outputs wait in registers for their correction
      if (zon) {
        if (d.cnt == K2_T) {                             // (forward: dir = +1)
#pragma unroll
          for (int j = 0; j < K2_C; ++j) y[j] = modal_step(cf, zsf(mine[j], true, mean, inv_sd), E);
        } else {
#pragma unroll
          for (int j = 0; j < K2_C; ++j) {
            const bool inside = tid * K2_C + j < d.cnt;  // zero padding stays zero under the z-score
            y[j] = modal_step(cf, inside ? zsf(mine[j], true, mean, inv_sd) : 0.0, E);
          }
        }
      } else if (dir > 0) {                              // the offsets are compile-time constants in either branch
#pragma unroll
        for (int j = 0; j < K2_C; ++j) y[j] = modal_step(cf, mine[j], E);
      } else {
#pragma unroll
        for (int j = 0; j < K2_C; ++j) y[j] = modal_step(cf, mine[-j], E);
      }
      K2_TICK(7)   // phase 1 recursion
      // phase 2: warp scan of chunk end states (Kogge-Stone with p^(C*2^k))
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        const int dd = 1 << kk;
        double o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __shfl_up_sync(AEW_FULL, E[i], dd);
        if (lane >= dd) cmul_acc(a.cPK + 4 * kk, o, E);
      }
      if (lane == 31) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s_W[4 * warp + i] = E[i];
      }
    }
    K2_TICK(1)   // recursion + warp scan (thread 0's own work)
    __syncthreads();
    K2_TICK(2)   // wait at the barrier
    K2_TICK_T(K2_PRODUCER, 0, true)

    // producer thread: the tile K2_AHEAD iterations from now
    if (tid == K2_PRODUCER) {
      place((it + K2_AHEAD) % K2_STAGES, true);
      K2_TICK_T(K2_PRODUCER, 12, false)
    }

    // =================== state entering my chunk, correction, store ===================
    // Every warp adds up the state entering it for itself (K2_NW - 1 products at most): S = the state entering the
    // TILE -- the zi-scaled state behind the odd extension at a segment's first tile, zero at the first tile a block
    // takes up for a run (a warm-up tile), else the state at the end of the tile before (s_chain: the block walks a
    // run of the segment in order, warm-up tiles first) -- times p^(32 C w), plus the totals of the warps before.
    if (filt) {
      const double* PWP = s_tab + K2_TAB_PWP;
      const double* Sp = d.k == 0 ? s_in : s_chain + 4 * (it & 1);
      const bool zero = d.k != 0 && (flags & K2_FRESH);
      const double S0[4] = {zero ? 0.0 : Sp[0], zero ? 0.0 : Sp[1], zero ? 0.0 : Sp[2], zero ? 0.0 : Sp[3]};
      double X[4] = {0.0, 0.0, 0.0, 0.0};
      cmul_acc(PWP + 4 * warp, S0, X);
#pragma unroll
      for (int v = 0; v < K2_NW - 1; ++v) {
        if (v < warp) {
          const double Wv[4] = {s_W[4 * v], s_W[4 * v + 1], s_W[4 * v + 2], s_W[4 * v + 3]};
          cmul_acc(PWP + 4 * (warp - 1 - v), Wv, X);
        }
      }
      if (tid == 0) {                                    // the state at the end of this tile, for the next one
        double Z[4] = {0.0, 0.0, 0.0, 0.0};
        cmul_acc(PWP + 4 * K2_NW, S0, Z);
#pragma unroll
        for (int v = 0; v < K2_NW; ++v) {
          const double Wv[4] = {s_W[4 * v], s_W[4 * v + 1], s_W[4 * v + 2], s_W[4 * v + 3]};
          cmul_acc(PWP + 4 * (K2_NW - 1 - v), Wv, Z);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s_chain[4 * ((it + 1) & 1) + i] = Z[i];
      }
      if (!(flags & K2_WARM)) {                          // (a warm-up tile leaves its end state and nothing else)
        // state entering my chunk = p^(C*lane) X + end state of the lane before me
        double s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = __shfl_up_sync(AEW_FULL, E[i], 1);
        if (lane == 0) { s[0] = s[1] = s[2] = s[3] = 0.0; }
        cmul_acc(s_tab + K2_TAB_PLANE + 4 * lane, X, s);
        // forward pass, last tile of a segment: state at the segment end -> right extension outputs
        if (!REV && d.n - d.t0 <= K2_T) {
          const int tl = (d.cnt - 1) / K2_C;             // thread holding the last valid sample
          if (tid == tl) {
            const int c = d.cnt - tl * K2_C;             // valid samples in its chunk, 1..C
            double z[4] = {s[0], s[1], s[2], s[3]};
            const double* xs = a.src + d.g0 + d.t0 + (long long)tl * K2_C;
            for (int j = 0; j < c; ++j) modal_step(cf, zsf(xs[j], zon, mean, inv_sd), z);
            const double* xe = a.src + d.g0 + d.n - 1;   // last sample of the segment
            const double xl = zsf(xe[0], zon, mean, inv_sd);
            double* ve = a.ws.vext + 16ll * d.seg;
            for (int m = 0; m < pad; ++m) ve[m] = modal_step(cf, 2.0 * xl - zsf(xe[-1 - m], zon, mean, inv_sd), z);
          }
        }
        // phase 3: add the zero-input response of the entering state, outputs to the tile buffer
        const double* G = a.cG;
        if (dir > 0) {
#pragma unroll
          for (int j = 0; j < K2_C; ++j)
            mine[j] = fma(G[4 * j + 3], s[3], fma(G[4 * j + 2], s[2], fma(G[4 * j + 1], s[1], fma(G[4 * j], s[0], y[j]))));
        } else {
#pragma unroll
          for (int j = 0; j < K2_C; ++j)
            mine[-j] = fma(G[4 * j + 3], s[3], fma(G[4 * j + 2], s[2], fma(G[4 * j + 1], s[1], fma(G[4 * j], s[0], y[j]))));
        }
      }
    }
    K2_TICK(3)   // entering states + correction
    const bool out = filt && !(flags & K2_WARM);
    if (out && d.bulk) fence_async_smem();               // my shared-memory writes -> visible to the bulk engine
    __syncthreads();
    K2_TICK(4)
    if (out) {
      if (d.bulk) {
        if (tid == K2_PRODUCER) bulk_store(a.dst + d.gstart, sm, K2_T * 8u);
      } else {
        for (int j = tid; j < d.cnt; j += K2_NT) {
          const long long p = d.t0 + j;
          a.dst[REV ? (d.g0 + d.n - 1 - p) : (d.g0 + p)] = sm[j];
        }
        if (tid == K2_PRODUCER) bulk_commit_empty();
      }
    } else if (tid == K2_PRODUCER) {
      bulk_commit_empty();                               // one bulk group per iteration keeps wait_group.read 1 exact
    }
    K2_TICK(5)   // store hand-off
  }
  if (tid == K2_PRODUCER) bulk_store_wait_all();
}

// SciPy's lfilter step (direct form II transposed), separate multiply and add.  ORD is a template parameter and every
// loop over the state is unrolled: with run-time indices b, a and z would live in local memory.
template <int ORD>
__device__ __forceinline__ double df2t_step_exact(const double (&b)[5], const double (&a)[5], double x, double (&z)[4]) {
  const double y = __dadd_rn(z[0], __dmul_rn(b[0], x));
#pragma unroll
  for (int k = 0; k < ORD - 1; ++k)
    z[k] = __dsub_rn(__dadd_rn(z[k + 1], __dmul_rn(x, b[k + 1])), __dmul_rn(y, a[k + 1]));
  z[ORD - 1] = __dsub_rn(__dmul_rn(x, b[ORD]), __dmul_rn(y, a[ORD]));
  return y;
}

// ------------------------------------------------------------ exact mode, low-latency form
// One WARP per segment: lane 0 walks the chain, all lanes load the next 64 samples, z-score them, form the products
// b[k] x into shared memory and store the outputs.  44 cycles per sample -- half the latency of the lane-per-segment
// form below, whose chain warp also carries the products and the z-score -- at 32 x its issue slots.  Used where the
// device has nothing else to do and the wait is what counts: the end-to-end feeder (PCIe-bound, the last chunk's
// pass is a tail of the call).
#define K2W_NW 1                        // one WARP per segment (the low-latency form), one per block.  Measured
#define K2W_NT (32 * K2W_NW)            // (profiles/README.md, "the end of a run"): 160-200 ms for the four passes of a
                                        // 1 h record with an SM sub-partition to itself; 400 ms with four such warps
                                        // in one block, one per sub-partition
#define K2W_HOG (40 * 1024)
#define K2W_BLK 64                      // samples per round: two per lane (halves the rounds' fixed cost)
// one pass over n samples: logical sample i is in[i] (REV: in[n - 1 - i]); lane 0 owns the state z
template <int ORD, bool REV>
__device__ __forceinline__ void exact_stream_warp(const double* in, double* out, long long n, const double* b,
                                             const double* a, double* z, bool zon, double mean, const ExactDiv& sd,
                                             double (*sp)[K2W_BLK + 2], double* sy, int lane) {
  double v0 = 0.0, v1 = 0.0;
  if (lane < n) v0 = in[REV ? n - 1 - lane : lane];
  if (lane + 32 < n) v1 = in[REV ? n - 1 - (lane + 32) : lane + 32];
  for (long long i0 = 0; i0 < n; i0 += K2W_BLK) {
    const long long j0 = i0 + K2W_BLK + lane, j1 = j0 + 32;   // the next round is in flight while lane 0 works
    double w0 = 0.0, w1 = 0.0;
    if (j0 < n) w0 = in[REV ? n - 1 - j0 : j0];
    if (j1 < n) w1 = in[REV ? n - 1 - j1 : j1];
    const double x0 = zs(v0, zon, mean, sd), x1 = zs(v1, zon, mean, sd);
#pragma unroll
    for (int k = 0; k <= ORD; ++k) { sp[k][lane] = __dmul_rn(b[k], x0); sp[k][lane + 32] = __dmul_rn(b[k], x1); }
    __syncwarp();
    if (lane == 0) {
      const int cnt = (int)(n - i0 < K2W_BLK ? n - i0 : K2W_BLK);
      // the products of the next two samples are fetched BEFORE this sample's dependent chain (DADD, DMUL, DADD)
      // is issued: an in-order warp would otherwise expose one shared-memory latency per sample
      double c[ORD + 1], c1[ORD + 1];
#pragma unroll
      for (int k = 0; k <= ORD; ++k) { c[k] = sp[k][0]; c1[k] = sp[k][1]; }
#pragma unroll 4
      for (int t = 0; t < cnt; ++t) {
        double c2[ORD + 1];
#pragma unroll
        for (int k = 0; k <= ORD; ++k) c2[k] = sp[k][t + 2];          // (the two pad columns keep this in bounds)
        const double y = __dadd_rn(z[0], c[0]);
#pragma unroll
        for (int k = 0; k < ORD - 1; ++k) z[k] = __dsub_rn(__dadd_rn(z[k + 1], c[k + 1]), __dmul_rn(y, a[k + 1]));
        z[ORD - 1] = __dsub_rn(c[ORD], __dmul_rn(y, a[ORD]));
        sy[t] = y;
#pragma unroll
        for (int k = 0; k <= ORD; ++k) { c[k] = c1[k]; c1[k] = c2[k]; }
      }
    }
    __syncwarp();
    if (i0 + lane < n) out[REV ? n - 1 - (i0 + lane) : i0 + lane] = sy[lane];
    if (i0 + lane + 32 < n) out[REV ? n - 1 - (i0 + lane + 32) : i0 + lane + 32] = sy[lane + 32];
    __syncwarp();
    v0 = w0; v1 = w1;
  }
}

template <int ORD>
__global__ void __launch_bounds__(K2W_NT) k2_exact_warp(PassArgs a) {
  __shared__ double sp_all[K2W_NW][ORD + 1][K2W_BLK + 2];
  __shared__ double sy_all[K2W_NW][K2W_BLK];
  const long long n_seg = a.ws.hdr[H_SEG_COUNT] < a.ws.max_seg ? a.ws.hdr[H_SEG_COUNT] : a.ws.max_seg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long seg = (long long)blockIdx.x * K2W_NW + warp;
  double (*sp)[K2W_BLK + 2] = sp_all[warp];
  double* sy = sy_all[warp];
  if (seg >= n_seg) return;                              // (warps are independent: only warp barriers below)
  const long long* sg = a.ws.segs + 4 * seg;
  const long long g0 = sg[0], n = sg[1], rec = sg[2];
  const bool zon = a.mean != nullptr;
  const double mean = zon ? a.mean[rec] : 0.0;
  const ExactDiv sd = make_exact_div(zon ? a.sd[rec] : 1.0);
  const int pad = (int)a.tab[K2_TAB_PAD];
  double b[5], ac[5], zi[4], z[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < 5; ++i) { b[i] = a.tab[K2_TAB_B + i]; ac[i] = a.tab[K2_TAB_A + i]; }
  for (int i = 0; i < 4; ++i) zi[i] = a.tab[K2_TAB_ZI + i];
  const double* xs = a.src + g0;
  double* out = a.passthrough + g0;
  double* tmp = a.dst + g0;
  if (n <= pad) {                                        // filters.py:50-51: too short to filter, copied through
    for (long long i = lane; i < n; i += 32) out[i] = zs(xs[i], zon, mean, sd);
    return;
  }
  double ve[K2_MAXPAD];
  const double x0 = zs(xs[0], zon, mean, sd), xl = zs(xs[n - 1], zon, mean, sd);
  if (lane == 0) {                                       // odd extension on the left, zi scaled by its first sample
    const double e0 = __dsub_rn(__dmul_rn(2.0, x0), zs(xs[pad], zon, mean, sd));
    for (int i = 0; i < ORD; ++i) z[i] = __dmul_rn(zi[i], e0);
    for (int e = 0; e < pad; ++e)
      df2t_step_exact<ORD>(b, ac, __dsub_rn(__dmul_rn(2.0, x0), zs(xs[pad - e], zon, mean, sd)), z);
  }
  exact_stream_warp<ORD, false>(xs, tmp, n, b, ac, z, zon, mean, sd, sp, sy, lane);
  if (lane == 0) {                                       // right extension, then the backward pass starts inside it
    for (int m = 0; m < pad; ++m)
      ve[m] = df2t_step_exact<ORD>(b, ac, __dsub_rn(__dmul_rn(2.0, xl), zs(xs[n - 2 - m], zon, mean, sd)), z);
    for (int i = 0; i < ORD; ++i) z[i] = __dmul_rn(zi[i], ve[pad - 1]);
    for (int e = 0; e < pad; ++e) df2t_step_exact<ORD>(b, ac, ve[pad - 1 - e], z);
  }
  __syncwarp();                                          // (orders the warp's stores to tmp before the reads below)
  exact_stream_warp<ORD, true>(tmp, out, n, b, ac, z, false, 0.0, sd, sp, sy, lane);
}

// ------------------------------------------------------------ forward + backward in one launch
// A filter whose memory is short (the 30 Hz low-pass: |p| <= 0.868, a state error is below 1e-18 after 300
// samples) needs no state from outside a tile's neighbourhood: a block takes a tile (K2_T samples) TOGETHER with a
// halo of 416 samples on either side, runs the forward pass over the K2_T + 832 samples from a zero state (the real
// initial state at a segment's first sample), reverses them in shared memory, runs the backward pass the same way
// (the real initial state at a segment's last sample, from the odd extension), and stores the middle.  One launch
// instead of two, 8 B read + 8 B written per sample instead of 32 (the halo re-reads come from L2), no dependency
// between tiles.  The 1 Hz high-pass (|p| = 0.995, ~8,000 samples of memory) stays with k2_pass.
#define K2F_NT (K2_NT + 64)             // warp 0 and the last warp hold the two halos of an interior tile
#define K2F_NW (K2F_NT / 32)
#define K2F_H (32 * K2_C)               // 416
#define K2F_R (K2F_NT * K2_C)           // = K2_T + 2 * K2F_H

// One pass over buf[0 .. K2F_R) in place: zero-state recursion per 13-sample chunk, scan of the chunk end states
// over the warp and over the ten warps, outputs corrected with the zero-input response of the entering state.
// S0 (shared memory, 4 doubles) is the state entering sample 0, used when `first`.  Returns the state entering the
// caller's chunk (needed for the right extension).
__device__ __forceinline__ void fused_pass(const PassArgs& a, const double* s_tab, double* buf, double* s_W,
                                           const double* S0, bool first, int tid, double s_out[4]) {
  const int lane = tid & 31, warp = tid >> 5;
  const Modal& cf = a.cf;
  double E[4] = {0.0, 0.0, 0.0, 0.0};
  double* mine = buf + tid * K2_C;
#pragma unroll
  for (int j = 0; j < K2_C; ++j) mine[j] = modal_step(cf, mine[j], E);
  const double Eown[4] = {E[0], E[1], E[2], E[3]};
#pragma unroll
  for (int kk = 0; kk < 5; ++kk) {                       // inclusive scan of the chunk end states over the warp
    const int dd = 1 << kk;
    double o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __shfl_up_sync(AEW_FULL, E[i], dd);
    if (lane >= dd) cmul_acc(a.cPK + 4 * kk, o, E);
  }
  if (lane == 31) {
#pragma unroll
    for (int i = 0; i < 4; ++i) s_W[4 * warp + i] = E[i];
  }
  __syncthreads();
  // every warp scans the ten warp totals for itself: state entering warp w = sum_{v < w} p^(32 C (w - 1 - v)) W_v
  const double* PWP = s_tab + K2_TAB_PWP;                // p^(32 C m), m = 0 .. K2_NW (= 8)
  double I[4] = {0.0, 0.0, 0.0, 0.0};
  if (lane < K2F_NW) {
#pragma unroll
    for (int i = 0; i < 4; ++i) I[i] = s_W[4 * lane + i];
  }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int dd = 1 << kk;
    double o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __shfl_up_sync(AEW_FULL, I[i], dd);
    if (lane >= dd && lane < K2F_NW) cmul_acc(PWP + 4 * dd, o, I);
  }
  double X[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) X[i] = __shfl_sync(AEW_FULL, I[i], warp > 0 ? warp - 1 : 0);
  if (warp == 0) { X[0] = X[1] = X[2] = X[3] = 0.0; }
  if (first) {                                           // + p^(32 C w) S0
    const double s0[4] = {S0[0], S0[1], S0[2], S0[3]};
    if (warp <= K2_NW) {
      cmul_acc(PWP + 4 * warp, s0, X);
    } else {                                             // p^(32 C 9) = p^(32 C 8) p^(32 C)
      const double* P8 = PWP + 4 * K2_NW;
      const double* P1 = PWP + 4;
      const double P9[4] = {P8[0] * P1[0] - P8[1] * P1[1], P8[0] * P1[1] + P8[1] * P1[0],
                            P8[2] * P1[2] - P8[3] * P1[3], P8[2] * P1[3] + P8[3] * P1[2]};
      cmul_acc(P9, s0, X);
    }
  }
  // state entering my chunk = p^(C lane) X + inclusive scan of the lane before me
  double sc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) sc[i] = __shfl_up_sync(AEW_FULL, E[i], 1);
  if (lane == 0) { sc[0] = sc[1] = sc[2] = sc[3] = 0.0; }
  cmul_acc(s_tab + K2_TAB_PLANE + 4 * lane, X, sc);
  (void)Eown;
  const double* G = a.cG;
#pragma unroll
  for (int j = 0; j < K2_C; ++j)
    mine[j] = fma(G[4 * j + 3], sc[3], fma(G[4 * j + 2], sc[2], fma(G[4 * j + 1], sc[1], fma(G[4 * j], sc[0], mine[j]))));
#pragma unroll
  for (int i = 0; i < 4; ++i) s_out[i] = sc[i];
}

__global__ void __launch_bounds__(K2F_NT, 3) k2_fused(PassArgs a) {
  AEW_DYN_SMEM(smem_raw);
  double* buf0 = (double*)smem_raw;                     // [K2F_R] forward pass, in place
  double* buf1 = buf0 + K2F_R;                          // [K2F_R] the reversed forward output, backward pass in place
  double* s_tab = buf1 + K2F_R;                         // [K2_TAB_LEN]
  double* s_W = s_tab + K2_TAB_LEN;                     // [K2F_NW][4]
  double* s_in = s_W + 4 * K2F_NW;                      // [4] state entering the region's first sample (forward)
  double* s_inb = s_in + 4;                             // [4] state entering its last sample (backward)
  double* s_ve = s_inb + 4;                             // [16] forward outputs over the right extension
  __shared__ long long s_tile;
  const int tid = threadIdx.x;
  for (int i = tid; i < K2_TAB_LEN; i += K2F_NT) s_tab[i] = a.tab[i];
  const int pad = (int)a.tab[K2_TAB_PAD];
  const Modal& cf = a.cf;
  long long* work = a.ws.hdr + H_WORK_FWD;
  long long n_tiles = a.ws.hdr[H_TILE_TOTAL];
  if (n_tiles > a.ws.max_tiles) n_tiles = a.ws.max_tiles;
  for (;;) {
    __syncthreads();                                     // (previous tile done with s_tile and the buffers)
    if (tid == 0) s_tile = (long long)atomicAdd((unsigned long long*)work, 1ull);
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= n_tiles) break;
    long long rw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) rw[i] = a.ws.tile_desc[4 * tile + i];
    const long long g0 = rw[0], n = rw[1], rec = rw[2] >> 32, k = rw[3];
    const bool zon = a.mean != nullptr;
    const double mean = zon ? a.mean[rec] : 0.0;
    const ExactDiv sd = make_exact_div(zon ? a.sd[rec] : 1.0);
    const double inv_sd = zon ? 1.0 / a.sd[rec] : 1.0;
    if (n <= pad) {                                      // filters.py:50-51: too short, copy through
      for (long long i = tid; i < n; i += K2F_NT) a.dst[g0 + i] = zs(a.src[g0 + i], zon, mean, sd);
      continue;
    }
    const long long t0 = k * K2_T;
    const long long remain = n - t0;
    const int cnt = remain < K2_T ? (int)remain : K2_T;
    const long long r0 = t0 > K2F_H ? t0 - K2F_H : 0;
    const long long r1e = t0 + cnt + K2F_H;
    const long long r1 = r1e < n ? r1e : n;
    const int len = (int)(r1 - r0);
    const double* xs = a.src + g0 + r0;
    for (int j = tid; j < K2F_R; j += K2F_NT) buf0[j] = j < len ? zsf(xs[j], zon, mean, inv_sd) : 0.0;
    if (r0 == 0 && tid == 0) {                           // the odd extension on the left, from SciPy's zi-scaled state
      double z[4];
      const double x0 = zsf(xs[0], zon, mean, inv_sd);
      modal_initial(s_tab, 2.0 * x0 - zsf(xs[pad], zon, mean, inv_sd), z);
      for (int e = 0; e < pad; ++e) modal_step(cf, 2.0 * x0 - zsf(xs[pad - e], zon, mean, inv_sd), z);
#pragma unroll
      for (int i = 0; i < 4; ++i) s_in[i] = z[i];
    }
    __syncthreads();
    double sc[4];
    fused_pass(a, s_tab, buf0, s_W, s_in, r0 == 0, tid, sc);
    if (r1 == n) {                                       // the segment ends in this region: right extension outputs
      const int tl = (len - 1) / K2_C;                   // thread holding the last valid sample
      if (tid == tl) {
        const int c = len - tl * K2_C;                   // valid samples in its chunk, 1..C
        double z[4] = {sc[0], sc[1], sc[2], sc[3]};
        const double* xc = xs + (long long)tl * K2_C;
        for (int j = 0; j < c; ++j) modal_step(cf, zsf(xc[j], zon, mean, inv_sd), z);
        const double* xe = a.src + g0 + n - 1;           // last sample of the segment
        const double xl = zsf(xe[0], zon, mean, inv_sd);
        for (int m = 0; m < pad; ++m) s_ve[m] = modal_step(cf, 2.0 * xl - zsf(xe[-1 - m], zon, mean, inv_sd), z);
        double zb[4];                                    // the backward pass starts inside the extension
        modal_initial(s_tab, s_ve[pad - 1], zb);
        for (int e = 0; e < pad; ++e) modal_step(cf, s_ve[pad - 1 - e], zb);
#pragma unroll
        for (int i = 0; i < 4; ++i) s_inb[i] = zb[i];
      }
    }
    __syncthreads();
    for (int j = tid; j < K2F_R; j += K2F_NT) buf1[j] = j < len ? buf0[len - 1 - j] : 0.0;
    __syncthreads();
    fused_pass(a, s_tab, buf1, s_W, s_inb, r1 == n, tid, sc);
    __syncthreads();
    double* out = a.dst + g0 + t0;
    const int off = len - 1 - (int)(t0 - r0);            // out[i] = buf1[off - i]
    for (int i = tid; i < cnt; i += K2F_NT) out[i] = buf1[off - i];
  }
}

// ------------------------------------------------------------ exact sequential mode
// SciPy's evaluation order (lfilter's direct form II transposed, one sample at a time, separate multiply and add):
// bit-identical to scipy.signal.filtfilt.  The recursion cannot be re-associated without changing the bits (two
// DF2T trajectories that start one ulp apart do not merge again: profiles/README.md), so every segment is ONE
// dependent chain -- DADD, DMUL, DADD per sample, 3 x 8 cycles, ~17 FP64 instructions in all -- and takes ~40 cycles
// per sample and pass whatever else happens: ~150 ms for the four passes of a 1 h record.  What the kernel can
// choose is how little of the machine that costs.  One LANE per segment: the 32 lanes of a warp walk 32 segments in
// lockstep, every FP64 instruction serves 32 chains, and a batch of 100 records is four warps -- they run next to
// the other kernels without taking anything noticeable from them (the earlier forms, one thread or one warp per
// segment, needed 32 x the issue slots for the same latency: measured 185 ms per batch with three passes in flight
// next to the window / filter kernels, against 66 ms of those kernels alone).  Memory stays coalesced: a second
// warp moves one 256-byte row per segment and round through shared memory (exact_pass below); a first
// lane-per-segment form with per-lane loads took 5 x longer (32 sectors per load instruction), a second one whose
// chain warp moved the rows itself 6 x (a lone in-order warp pays every shared-memory latency in full), and a
// three-warp one whose mover also formed the products b[k] x and the z-score for the chain warp 2 x (then the mover,
// again a lone in-order warp, is the slow one: 627 ms per pass against 310).
#define K2E_T 32                        // samples per segment and round (one 256-byte row per segment)
#define K2E_LD (K2E_T + 1)              // row stride in shared memory: lane l reads row l -> bank (l + t) mod 32
#ifndef K2E_NP
#define K2E_NP 1                        // (chain warp, mover warp) PAIRS per block.  One: with four pairs per block
#endif
#define K2E_NT (64 * K2E_NP)            // (one chain warp per SM sub-partition, 128 segments per SM instead of 32) a
                                        // pass took 460 ms instead of 313 and the calls next to it 73 ms instead of 70
#define K2E_HOG (190 * 1024)            // dynamic shared memory the block asks for and never touches: no block of the
                                        // window / filter / delineation kernels fits next to it, so the chain warp
                                        // keeps its SM sub-partition to itself.  The chain is a latency race: with
                                        // those kernels' warps on the same scheduler every one of its ~25
                                        // instructions per sample queues for an issue slot (measured: the pass went
                                        // from 315 ms alone to > 1 s next to them).  100 records are 4 blocks = 4 of
                                        // the 148 SMs.

// barrier of one pair (named barrier 1 + pair, 64 threads): the pairs of a block never wait for each other
#ifndef AEW_EMUL
__device__ __forceinline__ void aew_named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
#endif
__device__ __forceinline__ void pair_sync(int pair) { aew_named_sync(1 + pair, 64); }

struct alignas(16) K2ESeg { const double* in; double* out; long long n; long long pad_; };   // 32 bytes: two 16-byte loads
// shared memory of one pair: two input and two output tiles, the two segment tables, the longest segment
#define K2E_PAIR_BYTES (((4 * K2E_T * K2E_LD * 8 + 64 * 32 + 16) + 127) / 128 * 128)

// The mover warp's two jobs for one round.  Eight segments at a time: their table entries are loaded first (two
// 16-byte shared-memory loads each, all independent), then the eight copies are issued -- a lone warp runs in
// order, and with one load-then-use per row every shared-memory latency was paid in full (profiles/r2_k2_exact_simd.txt:
// 4,600 cycles per round, short scoreboard + wait).
template <bool REV>
__device__ __forceinline__ void exact_fetch(const K2ESeg* tab, long long r, double (*tile)[K2E_LD], int lane) {
  const long long i = r * K2E_T + lane;                   // logical index of my column in every row
#pragma unroll 1
  for (int g = 0; g < 32; g += 8) {
    const double* p[8];
    long long ns[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { p[u] = tab[g + u].in; ns[u] = tab[g + u].n; }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (i < ns[u]) {
        const double* src = p[u] + (REV ? ns[u] - 1 - i : i);
#ifdef AEW_EMUL
        tile[g + u][lane] = *src;
#else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&tile[g + u][lane])), "l"(src));
#endif
      }
    }
  }
#ifndef AEW_EMUL
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <bool REV>
__device__ __forceinline__ void exact_store(const K2ESeg* tab, long long r, const double (*tile)[K2E_LD], int lane) {
  const long long i = r * K2E_T + lane;
#pragma unroll 1
  for (int g = 0; g < 32; g += 8) {
    double* p[8];
    long long ns[8];
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { p[u] = tab[g + u].out; ns[u] = tab[g + u].n; v[u] = tile[g + u][lane]; }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (i < ns[u]) p[u][REV ? ns[u] - 1 - i : i] = v[u];
  }
}

// One pass of the block's 32 segments: logical sample i of a segment is in[i] (REV: in[n - 1 - i]).  Rounds of 32
// samples; in round r warp 0 walks row `lane` of tile r & 1 (every lane its own segment's chain) while warp 1
// fetches the rows of round r + 1 into the other tile (asynchronous copies, lane t takes column t: coalesced) and
// stores the outputs of round r - 1.  One block barrier per round.
template <int ORD, bool REV>
__device__ __forceinline__ void exact_pass(const K2ESeg* tab, long long n_max, double (*tin)[K2E_T][K2E_LD],
                                           double (*tout)[K2E_T][K2E_LD], const double (&b)[5], const double (&a)[5],
                                           double (&z)[4], long long n, bool zon, double mean, const ExactDiv& sd,
                                           int role, int pair, int lane) {
  const long long rounds = (n_max + K2E_T - 1) / K2E_T;
  if (role == 1 && rounds > 0) {
    exact_fetch<REV>(tab, 0, tin[0], lane);
#ifndef AEW_EMUL
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
  }
  pair_sync(pair);
  for (long long r = 0; r < rounds; ++r) {
    const int buf = (int)(r & 1);
    if (role == 0) {
      const long long i0 = r * K2E_T;
      if (i0 < n) {
        const int cnt = (int)(n - i0 < K2E_T ? n - i0 : K2E_T);
        const double* row = tin[buf][lane];
        double* orow = tout[buf][lane];
        double c = row[0];
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
          const double nx = row[t + 1];                    // (one pad column: in bounds; fetched before the chain)
          const double x = zs(c, zon, mean, sd);
          const double y = __dadd_rn(z[0], __dmul_rn(b[0], x));
#pragma unroll
          for (int k = 0; k < ORD - 1; ++k)
            z[k] = __dsub_rn(__dadd_rn(z[k + 1], __dmul_rn(x, b[k + 1])), __dmul_rn(y, a[k + 1]));
          z[ORD - 1] = __dsub_rn(__dmul_rn(x, b[ORD]), __dmul_rn(y, a[ORD]));
          orow[t] = y;
          c = nx;
        }
      }
    } else {
      if (r + 1 < rounds) exact_fetch<REV>(tab, r + 1, tin[buf ^ 1], lane);
      if (r > 0) exact_store<REV>(tab, r - 1, tout[buf ^ 1], lane);
#ifndef AEW_EMUL
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
    }
    pair_sync(pair);
  }
  if (role == 1 && rounds > 0) exact_store<REV>(tab, rounds - 1, tout[(rounds - 1) & 1], lane);
  pair_sync(pair);
}

template <int ORD>
__global__ void __launch_bounds__(K2E_NT) k2_exact(PassArgs a) {
  AEW_DYN_SMEM(smem_raw);                                // (the first K2E_PAIR_BYTES * K2E_NP bytes of the hog)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pair = warp % K2E_NP, role = warp / K2E_NP;  // role 0: the chains, role 1: the rows
  unsigned char* mine = smem_raw + (size_t)pair * K2E_PAIR_BYTES;
  double (*tin)[K2E_T][K2E_LD] = reinterpret_cast<double (*)[K2E_T][K2E_LD]>(mine);
  double (*tout)[K2E_T][K2E_LD] = reinterpret_cast<double (*)[K2E_T][K2E_LD]>(mine + 2 * K2E_T * K2E_LD * sizeof(double));
  K2ESeg* s_fwd = reinterpret_cast<K2ESeg*>(mine + 4 * K2E_T * K2E_LD * sizeof(double));   // src -> tmp
  K2ESeg* s_bwd = s_fwd + 32;                            // tmp -> out
  long long* s_nmax_p = reinterpret_cast<long long*>(s_bwd + 32);
  const long long n_seg = a.ws.hdr[H_SEG_COUNT] < a.ws.max_seg ? a.ws.hdr[H_SEG_COUNT] : a.ws.max_seg;
  const long long seg = ((long long)blockIdx.x * K2E_NP + pair) * 32 + lane;
  const bool have = seg < n_seg;
  const long long* sg = a.ws.segs + 4 * (have ? seg : 0);
  const long long g0 = have ? sg[0] : 0, rec = have ? sg[2] : 0;
  long long n = have ? sg[1] : 0;
  const bool zon = a.mean != nullptr;
  const double mean = zon ? a.mean[rec] : 0.0;
  const ExactDiv sd = make_exact_div(zon ? a.sd[rec] : 1.0);
  const int pad = (int)a.tab[K2_TAB_PAD];
  double b[5], ac[5], zi[4], z[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 5; ++i) { b[i] = a.tab[K2_TAB_B + i]; ac[i] = a.tab[K2_TAB_A + i]; }
#pragma unroll
  for (int i = 0; i < 4; ++i) zi[i] = a.tab[K2_TAB_ZI + i];
  const double* xs = a.src + g0;
  double* out = a.passthrough + g0;
  double* tmp = a.dst + g0;
  if (role == 0) {
    if (have && n <= pad) {                              // filters.py:50-51: too short to filter, copied through
      for (long long i = 0; i < n; ++i) out[i] = zs(xs[i], zon, mean, sd);
      n = 0;                                             // (takes no further part, but stays with its warp)
    }
    s_fwd[lane].in = xs; s_fwd[lane].out = tmp; s_fwd[lane].n = n;
    s_bwd[lane].in = tmp; s_bwd[lane].out = out; s_bwd[lane].n = n;
    const long long m = (long long)__reduce_max_sync(AEW_FULL, (unsigned)(n > 0x7fffffffll ? 0x7fffffffll : n));
    if (lane == 0) *s_nmax_p = m;
  } else {
    n = 0;
  }
  pair_sync(pair);
  const long long n_max = *s_nmax_p;
  double ve[K2_MAXPAD];
  double x0 = 0.0, xl = 0.0;
  if (role == 0 && n > 0) {                              // odd extension on the left, zi scaled by its first sample
    x0 = zs(xs[0], zon, mean, sd); xl = zs(xs[n - 1], zon, mean, sd);
    const double e0 = __dsub_rn(__dmul_rn(2.0, x0), zs(xs[pad], zon, mean, sd));
#pragma unroll
    for (int i = 0; i < ORD; ++i) z[i] = __dmul_rn(zi[i], e0);
    for (int e = 0; e < pad; ++e)
      df2t_step_exact<ORD>(b, ac, __dsub_rn(__dmul_rn(2.0, x0), zs(xs[pad - e], zon, mean, sd)), z);
  }
  exact_pass<ORD, false>(s_fwd, n_max, tin, tout, b, ac, z, n, zon, mean, sd, role, pair, lane);
  if (role == 0 && n > 0) {                              // right extension, then the backward pass starts inside it
    for (int m = 0; m < pad; ++m)
      ve[m] = df2t_step_exact<ORD>(b, ac, __dsub_rn(__dmul_rn(2.0, xl), zs(xs[n - 2 - m], zon, mean, sd)), z);
    const double vl = ve[pad - 1];
#pragma unroll
    for (int i = 0; i < ORD; ++i) z[i] = __dmul_rn(zi[i], vl);
    for (int e = 0; e < pad; ++e) df2t_step_exact<ORD>(b, ac, ve[pad - 1 - e], z);
  }
  __threadfence_block();                                 // the pair's stores to tmp before the copies that read them
  pair_sync(pair);
  exact_pass<ORD, true>(s_bwd, n_max, tin, tout, b, ac, z, n, false, 0.0, sd, role, pair, lane);
}

unsigned long long* g_dbg = nullptr;   // device counters, allocated once when AEW_K2_TIMING is set
inline int sm_count() {                 // of the CURRENT device (queried per call: devices may differ)
#ifdef AEW_EMUL
  return 2;
#else
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
#endif
}

}  // namespace

// ------------------------------------------------------------------ C ABI
extern "C" long long aew_filt_max_tiles(long long n_rec, long long total_samples, long long max_seg) {
  return total_samples / K2_T + n_rec + max_seg + 1;
}

extern "C" size_t aew_filt_workspace_bytes(long long n_rec, long long total_samples, long long max_seg) {
  return k2_ws_layout(nullptr, max_seg, aew_filt_max_tiles(n_rec, total_samples, max_seg), nullptr) + 128;
}

// Finds the non-NaN segments of every record and (if nan_fill != NULL) writes NaN
// to nan_fill at every NaN input position.  nonnan_count (device, optional, from
// aew_zscore_stats_f64) lets NaN-free records skip the scan.
extern "C" int aew_segments_find_f64(const double* x, const long long* d_rec_beg, const long long* d_rec_end,
                                     long long n_rec,
                                     long long total_samples, const long long* d_nonnan_count, double* nan_fill,
                                     void* ws, size_t ws_bytes, long long max_seg, void* stream) {
  if (!x || !d_rec_beg || !d_rec_end || !ws || n_rec < 0 || max_seg < 1) return AEW_ERR_ARG;
  if (ws_bytes < aew_filt_workspace_bytes(n_rec, total_samples, max_seg)) return AEW_ERR_WORKSPACE;
  if (n_rec > 0x7fffffffll || max_seg > 0x7fffffffll) return AEW_ERR_UNSUPPORTED;
  K2Ws w;
  k2_ws_layout((void*)align_up((size_t)ws, 128), max_seg, aew_filt_max_tiles(n_rec, total_samples, max_seg), &w);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(w.hdr, 0, H_SLOTS * sizeof(long long), st) != 0) return (int)cudaGetLastError();
  if (n_rec == 0) return AEW_OK;
  AEW_LAUNCH(k2_seg_find, (unsigned)n_rec, SEG_THREADS, 0, st, x, d_rec_beg, d_rec_end, d_nonnan_count, nan_fill, w);
  return (int)cudaGetLastError();
}

// 1 if the filter's memory is short enough for the one-launch form: |p|^416 < 1e-18 for every pole
extern "C" int aew_filt_fused_ok(const double* h_tab) {
  if (!h_tab) return 0;
  const double m1 = h_tab[K2_TAB_P + 0] * h_tab[K2_TAB_P + 0] + h_tab[K2_TAB_P + 1] * h_tab[K2_TAB_P + 1];
  const double m2 = h_tab[K2_TAB_P + 2] * h_tab[K2_TAB_P + 2] + h_tab[K2_TAB_P + 3] * h_tab[K2_TAB_P + 3];
  const double m = m1 > m2 ? m1 : m2;                    // |p|^2
  if (!(m > 0.0) || !(m < 1.0)) return m == 0.0 ? 1 : 0;
  return 0.5 * K2F_H * log(m) < log(1e-18) ? 1 : 0;
}

// filtfilt of every segment found by aew_segments_find_f64 (same ws).
//   src  input samples (flat), optionally z-scored on load with mean/sd per record
//   out  filtered output (may alias src); tmp: scratch of the same length
//   mode 0 = tiled scan (fast), 1 = exact sequential SciPy order (one lane per segment: least machine), 3 = the same
//   bits, one warp per segment (least latency), 2 = tiled, forward + backward in one launch (src -> out directly,
//   tmp unused, src != out; only where aew_filt_fused_ok(h_tab))
extern "C" int aew_filtfilt_f64(const double* src, const double* mean, const double* sd, double* out, double* tmp,
                                const double* d_tab, const double* h_tab, long long n_rec, long long total_samples,
                                void* ws, size_t ws_bytes, long long max_seg, int mode, void* stream) {
  if (!src || !out || !tmp || !d_tab || !h_tab || !ws || (mean && !sd)) return AEW_ERR_ARG;
  if (ws_bytes < aew_filt_workspace_bytes(n_rec, total_samples, max_seg)) return AEW_ERR_WORKSPACE;
  K2Ws w;
  k2_ws_layout((void*)align_up((size_t)ws, 128), max_seg, aew_filt_max_tiles(n_rec, total_samples, max_seg), &w);
  cudaStream_t st = (cudaStream_t)stream;
  PassArgs a;
  a.src = src; a.mean = mean; a.sd = sd; a.dst = tmp; a.passthrough = out; a.tab = d_tab; a.ws = w;
  a.bulk_ok = ((((size_t)src | (size_t)out | (size_t)tmp) & 15) == 0) ? 1 : 0;
  a.dbg = nullptr;
  a.run_tiles = 0;
  a.cf.d = h_tab[K2_TAB_D];
  a.cf.p1r = h_tab[K2_TAB_P + 0]; a.cf.p1i = h_tab[K2_TAB_P + 1]; a.cf.p2r = h_tab[K2_TAB_P + 2]; a.cf.p2i = h_tab[K2_TAB_P + 3];
  a.cf.g0 = h_tab[K2_TAB_GC + 0]; a.cf.g1 = h_tab[K2_TAB_GC + 1]; a.cf.g2 = h_tab[K2_TAB_GC + 2]; a.cf.g3 = h_tab[K2_TAB_GC + 3];
  for (int i = 0; i < 4 * K2_C; ++i) a.cG[i] = h_tab[K2_TAB_G + i];
  for (int i = 0; i < 20; ++i) a.cPK[i] = h_tab[K2_TAB_PK + i];
#ifndef AEW_EMUL
  if (getenv("AEW_K2_TIMING")) {
    if (!g_dbg) { cudaMalloc(&g_dbg, 16 * sizeof(unsigned long long)); cudaMemset(g_dbg, 0, 16 * sizeof(unsigned long long)); }
    a.dbg = g_dbg;
  }
#endif
  if (mode == 1) {                                       // one lane per segment (lanes beyond the count exit)
    const long long blocks = (max_seg + 32 * K2E_NP - 1) / (32 * K2E_NP);
    const char* hog_env = getenv("AEW_K2E_HOG_KB");      // measurement switch; default K2E_HOG
    size_t hog = hog_env ? (size_t)atoi(hog_env) * 1024 : (size_t)K2E_HOG;
    if (hog < (size_t)K2E_PAIR_BYTES * K2E_NP) hog = (size_t)K2E_PAIR_BYTES * K2E_NP;   // (the pairs' tiles live in it)
    switch ((int)h_tab[K2_TAB_ORDER]) {
      case 1: if (AEW_SET_MAX_SMEM(k2_exact<1>, K2E_HOG) != 0) return (int)cudaGetLastError();
              AEW_LAUNCH(k2_exact<1>, (unsigned)blocks, K2E_NT, hog, st, a); break;
      case 2: if (AEW_SET_MAX_SMEM(k2_exact<2>, K2E_HOG) != 0) return (int)cudaGetLastError();
              AEW_LAUNCH(k2_exact<2>, (unsigned)blocks, K2E_NT, hog, st, a); break;
      case 3: if (AEW_SET_MAX_SMEM(k2_exact<3>, K2E_HOG) != 0) return (int)cudaGetLastError();
              AEW_LAUNCH(k2_exact<3>, (unsigned)blocks, K2E_NT, hog, st, a); break;
      case 4: if (AEW_SET_MAX_SMEM(k2_exact<4>, K2E_HOG) != 0) return (int)cudaGetLastError();
              AEW_LAUNCH(k2_exact<4>, (unsigned)blocks, K2E_NT, hog, st, a); break;
      default: return AEW_ERR_UNSUPPORTED;
    }
    return (int)cudaGetLastError();
  }
  if (mode == 3) {                                       // exact, one warp per segment (low latency)
    const long long blocks = (max_seg + K2W_NW - 1) / K2W_NW;
    // K2W_HOG of dynamic shared memory that is never touched: enough that none of these blocks lands next to a
    // lane-per-segment block (K2E_HOG) -- as plain 32-thread blocks all 100 of a call piled onto the few SMs those
    // had emptied, a dozen chains per SM, 540 ms instead of 160 -- and little enough that they start wherever a
    // block of the window / filter kernels retires (measured at the end of a run, profiles/README.md: 40 KB 200 ms,
    // 116 KB or a whole SM 320 ms -- the 100 blocks then run in two waves)
    const char* hog_env = getenv("AEW_K2W_HOG_KB");      // measurement switch; default K2W_HOG
    const size_t hog = hog_env ? (size_t)atoi(hog_env) * 1024 : (size_t)K2W_HOG;
    switch ((int)h_tab[K2_TAB_ORDER]) {
      case 1: if (AEW_SET_MAX_SMEM(k2_exact_warp<1>, K2E_HOG) != 0) return (int)cudaGetLastError();
              AEW_LAUNCH(k2_exact_warp<1>, (unsigned)blocks, K2W_NT, hog, st, a); break;
      case 2: if (AEW_SET_MAX_SMEM(k2_exact_warp<2>, K2E_HOG) != 0) return (int)cudaGetLastError();
              AEW_LAUNCH(k2_exact_warp<2>, (unsigned)blocks, K2W_NT, hog, st, a); break;
      case 3: if (AEW_SET_MAX_SMEM(k2_exact_warp<3>, K2E_HOG) != 0) return (int)cudaGetLastError();
              AEW_LAUNCH(k2_exact_warp<3>, (unsigned)blocks, K2W_NT, hog, st, a); break;
      case 4: if (AEW_SET_MAX_SMEM(k2_exact_warp<4>, K2E_HOG) != 0) return (int)cudaGetLastError();
              AEW_LAUNCH(k2_exact_warp<4>, (unsigned)blocks, K2W_NT, hog, st, a); break;
      default: return AEW_ERR_UNSUPPORTED;
    }
    return (int)cudaGetLastError();
  }
  if (mode == 2) {                                       // forward + backward in one launch (short-memory filters)
    if (src == out) return AEW_ERR_ARG;
    if (!aew_filt_fused_ok(h_tab)) return AEW_ERR_UNSUPPORTED;
    const size_t fsmem = (size_t)(2 * K2F_R + K2_TAB_LEN + 4 * K2F_NW + 8 + 16) * sizeof(double);
    if (AEW_SET_MAX_SMEM(k2_fused, fsmem) != 0) return (int)cudaGetLastError();
    if (cudaMemsetAsync(w.hdr + H_WORK_FWD, 0, 2 * sizeof(long long), st) != 0) return (int)cudaGetLastError();
    a.dst = out;
    const long long cap3 = (long long)sm_count() * 3;
    const long long grid3 = w.max_tiles < cap3 ? w.max_tiles : cap3;
    AEW_LAUNCH(k2_fused, (unsigned)grid3, K2F_NT, fsmem, st, a);
    return (int)cudaGetLastError();
  }
  if (mode != 0) return AEW_ERR_ARG;
  const size_t smem = (size_t)(K2_STAGES * K2_T + K2_TAB_LEN + 4 * K2_NW + 12) * sizeof(double);
  // (per call: the attribute is per DEVICE)
  if (AEW_SET_MAX_SMEM(k2_pass<false>, smem) != 0 || AEW_SET_MAX_SMEM(k2_pass<true>, smem) != 0)
    return (int)cudaGetLastError();
  if (cudaMemsetAsync(w.hdr + H_WORK_FWD, 0, 2 * sizeof(long long), st) != 0) return (int)cudaGetLastError();
  // tiles per run: ~106,000 samples.  The warm-up in front of a run (`lookback` tiles: ~10,000 samples of the 1 Hz
  // high-pass) is then ~9 % more tiles without stores; measured on 1,000 one-hour cases the filter stage takes the
  // same time with runs of 64, 128 and 256 tiles and 4 % longer with 32, and shorter runs keep more blocks busy on
  // small batches (a single one-hour record is 17 runs).  (The value is part of the result's definition below
  // 1e-16: a run starts from a warm-up truncated at 1e-18.  AEW_K2_RUN is a measurement switch.)
  {
    const char* run_env = getenv("AEW_K2_RUN");
    a.run_tiles = run_env ? atoi(run_env) : K2_RUN;
    if (a.run_tiles < 1) a.run_tiles = 1;
  }
#ifndef AEW_EMUL
  if (getenv("AEW_K2_OCC")) {                            // measurement aid: resident blocks per SM the launch gets
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k2_pass<false>, K2_NT, smem);
    fprintf(stderr, "k2_pass: %d threads, %zu B dynamic shared memory, %d blocks per SM (planned %d)\n", K2_NT, smem, nb,
            K2_BPS);
  }
#endif
  const long long cap = (long long)sm_count() * K2_BPS;
  const long long grid = w.max_tiles < cap ? w.max_tiles : cap;
  AEW_LAUNCH(k2_pass<false>, (unsigned)grid, K2_NT, smem, st, a);
  PassArgs r = a;
  r.src = tmp; r.mean = nullptr; r.sd = nullptr; r.dst = out;
  AEW_LAUNCH(k2_pass<true>, (unsigned)grid, K2_NT, smem, st, r);
  return (int)cudaGetLastError();
}

// Reads back nothing: copies the 8-slot header (segment count, tile count, ...,
// status bits) into a caller-provided DEVICE buffer so the host can fetch it with
// its own (stream-ordered) copy.
extern "C" int aew_filt_header(void* ws, long long** d_header_out) {
  if (!ws || !d_header_out) return AEW_ERR_ARG;
  *d_header_out = (long long*)align_up((size_t)ws, 128);
  return AEW_OK;
}

// debugging aid (AEW_K2_TIMING=1): cycles thread 0 of every block spent in each phase of k2_pass, summed
extern "C" int aew_filt_timing_dump(unsigned long long* h_out8) {
#ifndef AEW_EMUL
  if (!g_dbg || !h_out8) return AEW_ERR_ARG;
  cudaDeviceSynchronize();
  cudaMemcpy(h_out8, g_dbg, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaMemset(g_dbg, 0, 16 * sizeof(unsigned long long));
  return AEW_OK;
#else
  (void)h_out8;
  return AEW_ERR_UNSUPPORTED;
#endif
}
