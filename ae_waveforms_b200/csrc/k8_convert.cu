// k8_convert.cu -- source samples (float32 or int16 + gain/offset) to the float64 the path computes in.
//
// The input side of the path: data_load_preprocess/load_record.py:7-37 hands
// process_case_id a float64 frame, but the recordings are stored narrower (VitalDB tracks are float32 or
// int16 with a gain and an offset).  Moving the narrow type over PCIe and widening it here halves or
// quarters the host->device bytes of the end-to-end path.  Values are exactly what NumPy's
// x.astype(np.float64) (float32) or v * gain + offset (int16; one multiplication, one addition, no FMA)
// would give; an int16 sentinel becomes NaN (a missing sample).
// HBM-bound: 4 + 8 (float32) or 2 + 8 (int16) bytes per sample; 16-byte loads and stores, grid-stride.
#include "aew_rt.h"
#include <math.h>

namespace {

__global__ void __launch_bounds__(256)
k8_from_f32(const float* __restrict__ src, long long n, double* __restrict__ dst) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = n >> 2;
#ifndef AEW_EMUL
  const float4* s4 = reinterpret_cast<const float4*>(src);
  double2* d2 = reinterpret_cast<double2*>(dst);
  for (long long i = t; i < n4; i += stride) {
    const float4 v = s4[i];
    double2 a, b;
    a.x = (double)v.x; a.y = (double)v.y; b.x = (double)v.z; b.y = (double)v.w;
    d2[2 * i] = a;
    d2[2 * i + 1] = b;
  }
#else
  for (long long i = t; i < n4; i += stride)
    for (int k = 0; k < 4; ++k) dst[4 * i + k] = (double)src[4 * i + k];
#endif
  for (long long i = 4 * n4 + t; i < n; i += stride) dst[i] = (double)src[i];
}

__global__ void __launch_bounds__(256)
k8_from_i16(const short* __restrict__ src, long long n, double gain, double offset, int use_sentinel, int sentinel,
            double* __restrict__ dst) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const long long n8 = n >> 3;
#ifndef AEW_EMUL
  const int4* s8 = reinterpret_cast<const int4*>(src);
  double2* d2 = reinterpret_cast<double2*>(dst);
  for (long long i = t; i < n8; i += stride) {
    const int4 q = s8[i];
    const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int lo = (short)(w[k] & 0xffff), hi = (short)((unsigned)w[k] >> 16);
      double2 o;
      o.x = (use_sentinel && lo == sentinel) ? nan : __dadd_rn(__dmul_rn((double)lo, gain), offset);
      o.y = (use_sentinel && hi == sentinel) ? nan : __dadd_rn(__dmul_rn((double)hi, gain), offset);
      d2[4 * i + k] = o;
    }
  }
#else
  for (long long i = t; i < n8; i += stride)
    for (int k = 0; k < 8; ++k) {
      const int v = src[8 * i + k];
      dst[8 * i + k] = (use_sentinel && v == sentinel) ? nan : __dadd_rn(__dmul_rn((double)v, gain), offset);
    }
#endif
  for (long long i = 8 * n8 + t; i < n; i += stride) {
    const int v = src[i];
    dst[i] = (use_sentinel && v == sentinel) ? nan : __dadd_rn(__dmul_rn((double)v, gain), offset);
  }
}

// Small tables (record bounds, window starts, np_info rows) fetched by the SMs straight from page-locked host
// memory: a cudaMemcpyAsync of them would queue on the host->device copy engine BEHIND the next chunk's
// 1.4 GB transfer (the engine serves its queue in order, whatever the stream), stalling the kernels that need them.
__global__ void __launch_bounds__(256)
k8_fetch_host(const unsigned char* __restrict__ src, long long n_bytes, unsigned char* __restrict__ dst) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n16 = n_bytes >> 4;
#ifndef AEW_EMUL
  const int4* s = reinterpret_cast<const int4*>(src);
  int4* d = reinterpret_cast<int4*>(dst);
  for (long long i = t; i < n16; i += stride) d[i] = s[i];
#else
  for (long long i = t; i < n16; i += stride)
    for (int k = 0; k < 16; ++k) dst[16 * i + k] = src[16 * i + k];
#endif
  for (long long i = 16 * n16 + t; i < n_bytes; i += stride) dst[i] = src[i];
}

}  // namespace

extern "C" int aew_fetch_host(const void* h_pinned, long long n_bytes, void* d_out, void* stream) {
  if (n_bytes < 0) return AEW_ERR_ARG;
  if (n_bytes == 0) return AEW_OK;
  if (!h_pinned || !d_out) return AEW_ERR_ARG;
  if (((uintptr_t)h_pinned & 15) || ((uintptr_t)d_out & 15)) return AEW_ERR_ARG;   // 16-byte accesses
  long long grid = (n_bytes + 256 * 16 - 1) / (256 * 16);
  if (grid > 148LL * 4) grid = 148LL * 4;
  AEW_LAUNCH(k8_fetch_host, (unsigned)grid, 256, 0, stream, (const unsigned char*)h_pinned, n_bytes,
             (unsigned char*)d_out);
  return (int)cudaGetLastError();
}

extern "C" int aew_widen_f64(const void* d_src, int src_type, long long n, double gain, double offset,
                             int use_sentinel, int sentinel, double* d_out, void* stream) {
  if (n < 0 || (src_type != AEW_SRC_F32 && src_type != AEW_SRC_I16)) return AEW_ERR_ARG;
  if (n == 0) return AEW_OK;
  if (!d_src || !d_out) return AEW_ERR_ARG;
  if (((uintptr_t)d_src & 15) || ((uintptr_t)d_out & 15)) return AEW_ERR_ARG;   // 16-byte accesses
  const long long per_block = 256LL * (src_type == AEW_SRC_F32 ? 4 : 8);
  long long grid = (n + per_block - 1) / per_block;
  if (grid > 148LL * 16) grid = 148LL * 16;
  if (src_type == AEW_SRC_F32)
    AEW_LAUNCH(k8_from_f32, (unsigned)grid, 256, 0, stream, (const float*)d_src, n, d_out);
  else
    AEW_LAUNCH(k8_from_i16, (unsigned)grid, 256, 0, stream, (const short*)d_src, n, gain, offset, use_sentinel,
               sentinel, d_out);
  return (int)cudaGetLastError();
}
