// cabi_misc.cu -- version and device checks of the C ABI (include/aewave.h).
#include "aew_rt.h"

extern "C" int aew_version(void) { return AEW_VERSION; }

extern "C" int aew_device_check(void) {
#ifdef AEW_EMUL
  return AEW_OK;
#else
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return AEW_ERR_NO_DEVICE;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return AEW_ERR_NO_DEVICE;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 ? AEW_OK : AEW_ERR_NO_DEVICE;
#endif
}

// The heads of many equally spaced host records in ONE strided host->device copy (cudaMemcpy2DAsync): the preview
// of the end-to-end feeder (batched.run_host_pipelined) -- a thousand separate 160 KB copies cost the host 10 ms to
// issue and the copy engine a per-copy overhead each.  Pitches and width in bytes.
extern "C" int aew_copy_heads_h2d(void* d_dst, size_t dst_pitch, const void* h_src, size_t src_pitch, size_t width,
                                  size_t rows, void* stream) {
#ifdef AEW_EMUL
  for (size_t r = 0; r < rows; ++r) memcpy((char*)d_dst + r * dst_pitch, (const char*)h_src + r * src_pitch, width);
  (void)stream;
  return AEW_OK;
#else
  if (!d_dst || !h_src || width > dst_pitch || width > src_pitch) return AEW_ERR_ARG;
  if (rows == 0 || width == 0) return AEW_OK;
  return (int)cudaMemcpy2DAsync(d_dst, dst_pitch, h_src, src_pitch, width, rows, cudaMemcpyHostToDevice,
                                (cudaStream_t)stream);
#endif
}
