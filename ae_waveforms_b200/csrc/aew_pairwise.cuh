// aew_pairwise.cuh -- NumPy's float64 pairwise summation order on the GPU.
//
// np.add.reduce over n contiguous doubles is a fixed binary tree (SURVEY.md
// appendix B.5, restated and pinned in oracle/np_pairwise.py):
//   n <= 128 : eight stride-8 partial sums, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
//              then the tail sequentially (n < 8: plain left-to-right from 0.0)
//   n  > 128 : split at h = n/2 rounded down to a multiple of 8, left + right.
// The kernels reproduce the tree, not just the value: a range is addressed as a
// "leaf slot" of the complete binary tree of depth D = pw_depth(n); a range that
// stops splitting early keeps itself as its left descendant and an EMPTY right
// descendant, so slot -> (offset, size) is a walk over the bits of the slot
// index and needs no table.  Eight lanes own the eight partial sums of a leaf
// (64-byte coalesced reads); the xor-butterfly over those lanes reproduces the
// combine order because IEEE addition is commutative.  All arithmetic uses
// explicit round-to-nearest intrinsics so FMA contraction cannot change it
// (reference semantics: mean/std of data_load_preprocess/filters.py:12 and
// np.std at data_load_preprocess/remove_samples.py:62).
#pragma once
#include "aew_rt.h"

#define AEW_PW_LEAF 128

__host__ __device__ inline int pw_depth(long long n) {
  int d = 0;
  long long s = n;
  while (s > AEW_PW_LEAF) {
    long long h = s / 2;
    h -= h % 8;
    s -= h;  // the right child is the larger one
    ++d;
  }
  return d;
}

// Descend `levels` levels from the node (off, size) along the bits of idx (most significant first).
__host__ __device__ inline void pw_walk(long long& off, long long& size, int levels, unsigned idx) {
  for (int lvl = levels - 1; lvl >= 0; --lvl) {
    const unsigned bit = (idx >> lvl) & 1u;
    if (size > AEW_PW_LEAF) {
      long long h = size / 2;
      h -= h % 8;
      if (bit) { off += h; size -= h; } else { size = h; }
    } else if (bit) {
      size = 0;
      return;
    }
  }
}

// (offset, size) of node `idx` at `depth` levels below the root of a range of n.
__host__ __device__ inline void pw_node(long long n, int depth, unsigned idx, long long& off, long long& size) {
  off = 0;
  size = n;
  pw_walk(off, size, depth, idx);
}

// Sum of f(off .. off+size) for one leaf (size <= 128) by the 8 lanes of a group
// (g = lane & 7).  Every lane of the warp must call it (shuffles); the result is
// returned on all 8 lanes of the group.  f(i, primary): `primary` is true for
// exactly one call per element (lets a functor count elements).
template <class F>
__device__ __forceinline__ double pw_leaf_sum8(F& f, long long off, int size, int g) {
  double v[16];
  const int n8 = size & ~7;
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = (size >= 8 && 8 * k < n8) ? f(off + 8 * k + g, true) : 0.0;
  double r = v[0];
#pragma unroll
  for (int k = 1; k < 16; ++k)
    if (8 * k < n8) r = __dadd_rn(r, v[k]);
  double t = __shfl_xor_sync(AEW_FULL, r, 1);
  r = __dadd_rn(r, t);
  t = __shfl_xor_sync(AEW_FULL, r, 2);
  r = __dadd_rn(r, t);
  t = __shfl_xor_sync(AEW_FULL, r, 4);
  r = __dadd_rn(r, t);
  int i = n8;
  if (size < 8) { r = 0.0; i = 0; }
  for (; i < size; ++i) r = __dadd_rn(r, f(off + i, g == 0));
  return r;
}

// Combine 2^levels slot values (vals/valid in shared memory) up the tree, in
// place, by ONE warp (all 32 lanes call).  Result in vals[0] / valid[0].
__device__ __forceinline__ void pw_warp_combine(double* vals, unsigned char* valid, int levels, int lane) {
  for (int lvl = levels - 1; lvl >= 0; --lvl) {
    const int cnt = 1 << lvl;
    for (int base = 0; base < cnt; base += 32) {
      const int i = base + lane;
      double a = 0.0, b = 0.0;
      unsigned char va = 0, vb = 0;
      if (i < cnt) { a = vals[2 * i]; b = vals[2 * i + 1]; va = valid[2 * i]; vb = valid[2 * i + 1]; }
      __syncwarp();
      if (i < cnt) {
        vals[i] = (va && vb) ? __dadd_rn(a, b) : a;
        valid[i] = va;
      }
      __syncwarp();
    }
  }
}

// Block-cooperative NumPy-order sum of f(i), i in [0, n), n small enough that
// 2^pw_depth(n) slots fit `vals`/`valid` (shared).  All threads of the block call
// it; blockDim.x must be a multiple of 32.  Returns the sum on every thread.
template <class F>
__device__ inline double pw_block_sum(F& f, int n, int depth, double* vals, unsigned char* valid) {
  const int nslots = 1 << depth;
  const int group = threadIdx.x >> 3, g = threadIdx.x & 7, ngroups = blockDim.x >> 3;
  for (int s0 = 0; s0 < nslots; s0 += ngroups) {
    const int slot = s0 + group;
    long long off = 0, size = 0;
    if (slot < nslots) pw_node(n, depth, (unsigned)slot, off, size);
    const double v = pw_leaf_sum8(f, off, (int)size, g);
    if (slot < nslots && g == 0) { vals[slot] = v; valid[slot] = size > 0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) pw_warp_combine(vals, valid, depth, threadIdx.x);
  __syncthreads();
  const double r = vals[0];
  __syncthreads();
  return r;
}
