// aew_bulk.cuh -- 1-D bulk asynchronous copies (TMA: cp.async.bulk) and their mbarrier, shared by the filter
// (k2_filtfilt.cu: tile ring) and the window kernel (k3_windows.cu: parking / restoring the z-scored window).
// The emulation build (-DAEW_EMUL, tests only) copies synchronously.
#pragma once
#include <string.h>
#include "aew_rt.h"

#ifdef AEW_EMUL
__device__ __forceinline__ void mbar_init(unsigned long long*, int) {}
__device__ __forceinline__ bool mbar_wait(unsigned long long*, unsigned) { return true; }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulk_store(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulk_store_wait_read(int) {}
__device__ __forceinline__ void bulk_commit_empty() {}
__device__ __forceinline__ void bulk_store_wait_all() {}
__device__ __forceinline__ void fence_async_smem() {}
#else
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait: returns false if the phase never completed (reported through the status word)
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity) {
  for (unsigned tries = 0; tries < (1u << 22); ++tries) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
// global -> shared, completion counted in bytes on the mbarrier (arms it first)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_commit_empty() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_read(int pending) {
  if (pending == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif
