// aew_emul.h -- host emulation of the small CUDA subset the kernels use.
//
// DEVELOPMENT / TEST INFRASTRUCTURE ONLY (compiled with -DAEW_EMUL by
// tests/emul/Makefile into libaewave_emul.so, which the CPU test-suite loads to
// run the *kernel sources themselves* on host threads: one std::thread per CUDA
// thread of a block, blocks executed one after another).  The product library
// (libaewave.so, nvcc, sm_100a) never includes this file and has no CPU path.
//
// Supported: threadIdx/blockIdx/blockDim/gridDim, static and dynamic shared
// memory (blocks run sequentially, so `static` storage is block-shared),
// __syncthreads(_or), __syncwarp, __shfl_*_sync, __ballot_sync, __any_sync,
// atomics, fences, the arithmetic intrinsics with explicit rounding.
// Rule the kernels obey so that this works: warp collectives are executed by
// all live lanes of a warp, block barriers by all live threads of a block.
#pragma once
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
struct alignas(16) double2 { double x, y; };
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__ __restrict
#define __grid_constant__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef int cudaError_t;
typedef void* cudaStream_t;
static const cudaError_t cudaSuccess = 0;
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
inline cudaError_t cudaMemcpyAsyncH2D(void* d, const void* s, size_t n, cudaStream_t) { memcpy(d, s, n); return 0; }

namespace aew_emul {

class Barrier {
 public:
  void reset(int n) { live_ = n; waiting_ = 0; gen_ = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    unsigned g = gen_;
    if (++waiting_ >= live_) { waiting_ = 0; ++gen_; cv_.notify_all(); return; }
    cv_.wait(lk, [&] { return gen_ != g; });
  }
  void leave() {  // a thread that exits no longer takes part
    std::unique_lock<std::mutex> lk(m_);
    --live_;
    if (live_ > 0 && waiting_ >= live_) { waiting_ = 0; ++gen_; cv_.notify_all(); }
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  int live_ = 0, waiting_ = 0;
  unsigned gen_ = 0;
};

// named barrier (bar.sync / bar.arrive id, count): completes when `count` threads have arrived either way
class NamedBarrier {
 public:
  void arrive(int count) {
    std::unique_lock<std::mutex> lk(m_);
    if (++n_ >= count) { n_ = 0; ++gen_; cv_.notify_all(); }
  }
  void sync(int count) {
    std::unique_lock<std::mutex> lk(m_);
    const unsigned g = gen_;
    if (++n_ >= count) { n_ = 0; ++gen_; cv_.notify_all(); return; }
    cv_.wait(lk, [&] { return gen_ != g; });
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_ = 0;
  unsigned gen_ = 0;
};

struct WarpCtx {
  Barrier bar;
  unsigned long long slot[32];
  unsigned ballot_bits;
};

struct BlockCtx {
  Barrier bar;
  NamedBarrier named[16];
  std::vector<WarpCtx> warps;
  std::atomic<int> or_acc{0};
  std::vector<unsigned char> dyn;
};

inline BlockCtx*& block_ctx() { static BlockCtx* c = nullptr; return c; }
struct ThreadCtx { dim3 tid, bid, bdim, gdim; int lane, warp; };
inline ThreadCtx& tctx() { static thread_local ThreadCtx t; return t; }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F&& body) {
  const unsigned nthr = block.x * block.y * block.z;
  BlockCtx ctx;
  ctx.dyn.assign(smem + 64, 0xCD);
  ctx.warps = std::vector<WarpCtx>((nthr + 31) / 32);
  block_ctx() = &ctx;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        ctx.bar.reset((int)nthr);
        for (size_t w = 0; w < ctx.warps.size(); ++w) {
          int lanes = (int)std::min<unsigned>(32u, nthr - (unsigned)w * 32u);
          ctx.warps[w].bar.reset(lanes);
        }
        std::vector<std::thread> pool;
        pool.reserve(nthr);
        for (unsigned t = 0; t < nthr; ++t) {
          pool.emplace_back([&, t] {
            ThreadCtx& c = tctx();
            c.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            c.bid = dim3(bx, by, bz);
            c.bdim = block; c.gdim = grid;
            c.lane = (int)(t % 32); c.warp = (int)(t / 32);
            body();
            ctx.warps[c.warp].bar.leave();
            ctx.bar.leave();
          });
        }
        for (auto& th : pool) th.join();
      }
  block_ctx() = nullptr;
}

inline unsigned char* dyn_smem() {
  uintptr_t p = (uintptr_t)block_ctx()->dyn.data();
  return (unsigned char*)((p + 63) & ~(uintptr_t)63);
}
}  // namespace aew_emul

#define threadIdx (aew_emul::tctx().tid)
#define blockIdx (aew_emul::tctx().bid)
#define blockDim (aew_emul::tctx().bdim)
#define gridDim (aew_emul::tctx().gdim)

inline void __syncthreads() { aew_emul::block_ctx()->bar.wait(); }
inline int __syncthreads_or(int pred) {
  auto* b = aew_emul::block_ctx();
  if (pred) b->or_acc.store(1);
  b->bar.wait();
  int r = b->or_acc.load();
  b->bar.wait();
  if (aew_emul::tctx().tid.x == 0 && aew_emul::tctx().tid.y == 0) b->or_acc.store(0);
  b->bar.wait();
  return r;
}
inline int __syncthreads_and(int pred) { return !__syncthreads_or(!pred); }
inline void aew_named_sync(int id, int count) { aew_emul::block_ctx()->named[id].sync(count); }
inline void aew_named_arrive(int id, int count) { aew_emul::block_ctx()->named[id].arrive(count); }
inline void __syncwarp(unsigned = 0xffffffffu) {
  aew_emul::block_ctx()->warps[aew_emul::tctx().warp].bar.wait();
}

template <class T>
inline T aew_emul_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  auto& w = aew_emul::block_ctx()->warps[aew_emul::tctx().warp];
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w.slot[aew_emul::tctx().lane] = raw;
  w.bar.wait();
  unsigned long long got = w.slot[src_lane & 31];
  w.bar.wait();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int lane = aew_emul::tctx().lane;
  int base = lane & ~(width - 1);
  return aew_emul_exchange(v, base + (src & (width - 1)));
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  (void)width;
  return aew_emul_exchange(v, aew_emul::tctx().lane ^ m);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  int lane = aew_emul::tctx().lane;
  int src = lane - (int)d;
  if (src < (lane & ~(width - 1))) src = lane;
  return aew_emul_exchange(v, src);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  int lane = aew_emul::tctx().lane;
  int src = lane + (int)d;
  if (src > (lane | (width - 1))) src = lane;
  return aew_emul_exchange(v, src);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  auto& w = aew_emul::block_ctx()->warps[aew_emul::tctx().warp];
  w.slot[aew_emul::tctx().lane] = pred ? 1ull : 0ull;
  w.bar.wait();
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) r |= (unsigned)(w.slot[l] & 1ull) << l;
  w.bar.wait();
  return r;
}
inline unsigned aew_emul_reduce_u32(unsigned v, int op) {
  auto& w = aew_emul::block_ctx()->warps[aew_emul::tctx().warp];
  w.slot[aew_emul::tctx().lane] = v;
  w.bar.wait();
  unsigned r = (unsigned)w.slot[0];
  for (int l = 1; l < 32; ++l) {
    const unsigned o = (unsigned)w.slot[l];
    r = op == 0 ? (o < r ? o : r) : (op == 1 ? (o > r ? o : r) : r + o);
  }
  w.bar.wait();
  return r;
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) { return aew_emul_reduce_u32(v, 0); }
inline unsigned __reduce_max_sync(unsigned, unsigned v) { return aew_emul_reduce_u32(v, 1); }
inline unsigned __reduce_add_sync(unsigned, unsigned v) { return aew_emul_reduce_u32(v, 2); }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

// ---- atomics / fences
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline double atomicAdd(double* p, double v) {
  unsigned long long* q = (unsigned long long*)p;
  unsigned long long old = __atomic_load_n(q, __ATOMIC_SEQ_CST), nw;
  double o;
  do { memcpy(&o, &old, 8); double n = o + v; memcpy(&nw, &n, 8); }
  while (!__atomic_compare_exchange_n(q, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return o;
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }

// ---- arithmetic with explicit rounding (the emulation build uses -ffp-contract=off)
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __drcp_rn(double a) { volatile double r = 1.0 / a; return r; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline int __double2hiint(double d) { long long r; memcpy(&r, &d, 8); return (int)(r >> 32); }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
template <class T> inline T __ldg(const T* p) { return *p; }
using std::isnan;
using std::isinf;
inline long long atomicMin(long long* p, long long v) {
  long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
inline long long atomicMax(long long* p, long long v) {
  long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
