// aew_rt.h -- one include for every kernel source.
//
// Product build (nvcc, sm_100a): CUDA runtime + the launch macro below.
// Emulation build (-DAEW_EMUL, host g++; tests only): aew_emul.h provides the
// same names on host threads so the kernel sources can be exercised on a box
// without a GPU.  There is no other difference between the two builds.
#pragma once
#include <stdint.h>

#ifdef AEW_EMUL
#include "aew_emul.h"
#define AEW_LAUNCH(kernel, grid, block, smem, stream, ...) \
  aew_emul::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })
#define AEW_DYN_SMEM(name) unsigned char* name = aew_emul::dyn_smem()
#define AEW_SET_MAX_SMEM(kernel, bytes) 0
#else
#include <cuda_runtime.h>
#define AEW_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define AEW_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define AEW_SET_MAX_SMEM(kernel, bytes) \
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif

// ---- status codes shared by device and host (returned through the C ABI)
enum {
  AEW_OK = 0,
  AEW_ERR_ARG = -1,         // null pointer / negative size / unsupported parameter
  AEW_ERR_WORKSPACE = -2,   // workspace too small
  AEW_ERR_UNSUPPORTED = -3, // e.g. filter order > 4, window longer than the kernel supports
  AEW_ERR_NO_DEVICE = -4,   // no CUDA device / wrong architecture
  // bits of the device-side status word (read back by the caller after the stream drains)
  AEW_DEV_SEG_OVERFLOW = 1,   // more NaN-delimited segments than the workspace was sized for
  AEW_DEV_TILE_OVERFLOW = 2,  // more filter tiles than the workspace was sized for
  AEW_DEV_LOOKBACK_TIMEOUT = 4, // a tile waited too long for its predecessor (should never happen)
  AEW_DEV_RUN_OVERFLOW = 8,   // more QRS runs in a window than the kernel's table holds
};

#define AEW_WARP 32
#define AEW_FULL 0xffffffffu
