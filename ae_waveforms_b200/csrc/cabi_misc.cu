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
