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

// ---- status codes shared by device and host: include/aewave.h
#include "../../include/aewave.h"

#define AEW_WARP 32
#define AEW_FULL 0xffffffffu
