// common.cuh -- shared helpers for the sm_100a kernels of libenerf_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/enerf_b200.h"

namespace enerf {

void set_error(const char* fmt, ...);

#define ENERF_REQUIRE(cond, code, ...)  \
  do {                                  \
    if (!(cond)) {                      \
      enerf::set_error(__VA_ARGS__);    \
      return (code);                    \
    }                                   \
  } while (0)

#define ENERF_CHECK_LAUNCH(name)                                                       \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      enerf::set_error("%s: launch failed: %s", (name), cudaGetErrorString(e__));      \
      return ENERF_ECUDA;                                                              \
    }                                                                                  \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are PER DEVICE: a process that
// drives several GPUs (net.to('cuda:1'), one thread per device) must set / query them on each.
constexpr int ENERF_MAX_DEVICES = 64;
struct PerDeviceSize {
  size_t v[ENERF_MAX_DEVICES] = {};
  size_t& cur() {
    int d = 0;
    cudaGetDevice(&d);
    return v[(d >= 0 && d < ENERF_MAX_DEVICES) ? d : 0];
  }
};
static inline int device_sm_count() {
  static int n[ENERF_MAX_DEVICES] = {};
  int d = 0;
  cudaGetDevice(&d);
  if (d < 0 || d >= ENERF_MAX_DEVICES) d = 0;
  if (n[d] == 0) cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d);
  return n[d];
}
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// n / d for a run-time divisor d >= 1 and n < 2^31 as multiply-high + add + shift (Granlund-Montgomery): q = (umulhi(n, m) + n) >> l with
// l = ceil(log2 d), m = floor(2^32 (2^l - d) / d) + 1.  An unsigned division by a run-time value costs ~20 instructions.
struct FastDiv {
  unsigned m, l;
  static FastDiv make(unsigned d) {
    FastDiv f;
    f.l = 0;
    while ((1ull << f.l) < d) ++f.l;
    f.m = (unsigned)((((1ull << f.l) - d) << 32) / d + 1);
    return f;
  }
  __host__ __device__ __forceinline__ unsigned div(unsigned n) const {
#ifdef __CUDA_ARCH__
    return (__umulhi(n, m) + n) >> l;
#else
    return (unsigned)((((unsigned long long)n * m) >> 32) + n) >> l;
#endif
  }
};

// torch.linspace(0, 1, steps) element i, same two-sided formula as ATen's RangeFactories kernel
// (step = 1/(steps-1); lower half start + step*i, upper half end - step*(steps-1-i)).
__host__ __device__ __forceinline__ float linspace01(int i, int steps) {
  if (steps <= 1) return 0.f;
  const float step = 1.0f / (float)(steps - 1);
  return (i < steps / 2) ? step * (float)i : 1.0f - step * (float)(steps - 1 - i);
}

// F.interpolate(mode='bilinear', align_corners=True) of a single-channel map (hi,wi) evaluated at
// output pixel (y,x) of an (ho,wo) grid; same lambda construction as ATen's upsample_bilinear2d.
__device__ __forceinline__ float bilinear_ac(const float* __restrict__ p, int hi, int wi, int ho, int wo, int y, int x) {
  const float rh = (ho > 1) ? (float)(hi - 1) / (float)(ho - 1) : 0.f;
  const float rw = (wo > 1) ? (float)(wi - 1) / (float)(wo - 1) : 0.f;
  const float h1r = rh * (float)y;
  const int h1 = (int)h1r;
  const int h1p = (h1 < hi - 1) ? 1 : 0;
  const float h1l = h1r - (float)h1, h0l = 1.f - h1l;
  const float w1r = rw * (float)x;
  const int w1 = (int)w1r;
  const int w1p = (w1 < wi - 1) ? 1 : 0;
  const float w1l = w1r - (float)w1, w0l = 1.f - w1l;
  const float* r0 = p + (size_t)h1 * wi + w1;
  const float* r1 = r0 + (size_t)h1p * wi;
  return h0l * (w0l * __ldg(r0) + w1l * __ldg(r0 + w1p)) + h1l * (w0l * __ldg(r1) + w1l * __ldg(r1 + w1p));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace enerf
