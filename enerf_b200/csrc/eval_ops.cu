// eval_ops.cu -- consumers of the rendered frame that the reference runs on the CPU after a full
// fp32 D2H copy:
//   * PSNR  (/root/reference/lib/evaluators/enerf.py:45-71: skimage peak_signal_noise_ratio over the
//            pixels with msk >= 1, data_range 1  ==  10 log10(1 / mean((gt - pred)^2)))
//   * uint8 frame pack for display / video (/root/reference/gui_human.py:88-91: img*255 -> uint8,
//            vertical flip;  lib/evaluators/enerf.py:63: (img*255.).astype(np.uint8))
// Both are single-pass HBM-bound kernels; the frame then leaves the GPU as 16 bytes (PSNR sums) or
// 1 byte per channel instead of 4.
#include "common.cuh"

namespace enerf {

// acc[0] += sum over selected pixels and 3 channels of (pred - gt)^2 ; acc[1] += number of values
__global__ void __launch_bounds__(256) psnr_sse_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                       const unsigned char* __restrict__ mask, int mask_elem, long long n_pix,
                                                       double* __restrict__ acc) {
  double s = 0.0, c = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (long long)gridDim.x * blockDim.x) {
    bool on = true;
    if (mask) {
      on = false;
      for (int b = 0; b < mask_elem; ++b) on |= mask[i * mask_elem + b] != 0;
    }
    if (on) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float d = __ldg(pred + i * 3 + k) - __ldg(gt + i * 3 + k);
        s += (double)(d * d);
      }
      c += 3.0;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    c += __shfl_xor_sync(0xffffffffu, c, off);
  }
  __shared__ double ws[8], wc[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) ws[warp] = s, wc[warp] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0, tc = 0;
    for (int w = 0; w < 8; ++w) ts += ws[w], tc += wc[w];
    atomicAdd(acc, ts);
    atomicAdd(acc + 1, tc);
  }
}

// out (H,W,3) uint8 = trunc(clamp(rgb,0,1) * 255), rows optionally flipped (GL texture order)
__global__ void pack_rgb8_kernel(const float* __restrict__ rgb, int H, int W, int flip, unsigned char* __restrict__ out) {
  const long long total = (long long)H * W * 3;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % 3);
    const long long pix = t / 3;
    const int x = (int)(pix % W), y = (int)(pix / W);
    const float v = fminf(fmaxf(__ldg(rgb + t), 0.f), 1.f) * 255.f;
    const int yy = flip ? H - 1 - y : y;
    out[((size_t)yy * W + x) * 3 + c] = (unsigned char)v;
  }
}

}  // namespace enerf

extern "C" int enerf_psnr_accumulate(const float* pred, const float* gt, const void* mask, int mask_elem_size, long long n_pixels,
                                     double* acc, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(pred && gt && acc, ENERF_EINVAL, "psnr_accumulate: null pointer");
  if (n_pixels <= 0) return ENERF_OK;
  const int blocks = (int)((n_pixels + 255) / 256 < 592 ? (n_pixels + 255) / 256 : 592);
  psnr_sse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, gt, reinterpret_cast<const unsigned char*>(mask), mask_elem_size, n_pixels, acc);
  ENERF_CHECK_LAUNCH("psnr_accumulate");
  return ENERF_OK;
}

extern "C" int enerf_pack_rgb8(const float* rgb, int H, int W, int flip_vertical, unsigned char* out, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(rgb && out && H > 0 && W > 0, ENERF_EINVAL, "pack_rgb8: bad arguments");
  const long long total = (long long)H * W * 3;
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  pack_rgb8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rgb, H, W, flip_vertical, out);
  ENERF_CHECK_LAUNCH("pack_rgb8");
  return ENERF_OK;
}
