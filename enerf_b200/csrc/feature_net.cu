// feature_net.cu -- FeatureNet.forward (/root/reference/lib/networks/enerf/feature_net.py:27-36)
// as 11 launches of hand-written kernels: 8 direct convolutions (conv.cuh) + toplayer 1x1 +
// two fused {lateral 1x1 conv + bilinear x2 up-sample (align_corners) + add} kernels
// (feature_net.py:24-25,32-33).  The S source images ride on the depth axis of the conv template.
#include "conv.cuh"
#include "tc_conv.cuh"

namespace enerf {

//            CIN COUT KD KH ST  TZ TY  TX PY PX COT CCH  RELU  PLANAR
using Conv00 = ConvTraits<3, 8, 1, 3, 1, 1, 32, 32, 2, 4, 8, 3, true, true>;
using Conv01 = ConvTraits<8, 8, 1, 3, 1, 1, 32, 32, 2, 4, 8, 8, true, false>;
using Conv10 = ConvTraits<8, 16, 1, 5, 2, 1, 16, 32, 1, 4, 8, 8, true, false>;
using Conv11 = ConvTraits<16, 16, 1, 3, 1, 1, 32, 32, 2, 4, 8, 8, true, false>;
using Conv20 = ConvTraits<16, 32, 1, 5, 2, 1, 8, 32, 1, 4, 8, 8, true, false>;
using Conv21 = ConvTraits<32, 32, 1, 3, 1, 1, 16, 32, 2, 4, 8, 8, true, false>;
using Top = ConvTraits<32, 32, 1, 1, 1, 1, 16, 32, 2, 4, 8, 8, false, false>;
using Smooth1 = ConvTraits<32, 16, 1, 3, 1, 1, 32, 32, 2, 4, 8, 8, false, false>;
using Smooth0 = ConvTraits<32, 8, 1, 3, 1, 1, 32, 32, 2, 4, 8, 8, false, false>;

// out (N,H,W,32) = bias + W[CIN][32] . lat_in (N,H,W,CIN)  +  bilinear_x2(up_in (N,H/2,W/2,32))
// thread = (pixel, g): output channels [4g, 4g+4) and [16+4g, 16+4g+4), i.e. float4 slices g and g+4 of the
// 32-channel record -- the four lanes of a pixel then read / write 64 contiguous bytes per instruction
// (two full sectors) instead of four half-used ones: the kernel is L1-throughput bound (ncu: l1tex 82 %)
template <int CIN>
__global__ void __launch_bounds__(256) lateral_upadd_kernel(const float* __restrict__ lat_in, const float* __restrict__ w,
                                                            const float* __restrict__ bias, const float* __restrict__ up_in,
                                                            float* __restrict__ out, int N, int H, int W) {
  __shared__ __align__(16) float w_s[CIN * 32];
  __shared__ float b_s[32];
  for (int e = threadIdx.x; e < CIN * 32; e += blockDim.x) w_s[e] = w[e];
  if (threadIdx.x < 32) b_s[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int total = N * H * W * 4;          // 32-bit index math (launcher checks the range): 64-bit div/mod per thread
  const int hi = H / 2, wi = W / 2;         // was 2/3 of this kernel's instructions
  const float rh = (H > 1) ? (float)(hi - 1) / (float)(H - 1) : 0.f;
  const float rw = (W > 1) ? (float)(wi - 1) / (float)(W - 1) : 0.f;
  // the shared-memory weight reads were a third of the L1 data-pipe wavefronts of the full-resolution
  // instance (ncu): with 4 pixels per thread (the launch's grid-stride) the CIN = 8 slice is read once
  float4 wreg[CIN == 8 ? 16 : 1];
  if constexpr (CIN == 8) {
    const int g0 = threadIdx.x & 3;        // == t & 3 for every t of this thread (the stride is a multiple of 4)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      wreg[2 * k] = *reinterpret_cast<const float4*>(w_s + k * 32 + g0 * 4);
      wreg[2 * k + 1] = *reinterpret_cast<const float4*>(w_s + k * 32 + 16 + g0 * 4);
    }
  }
  // CIN = 16 (lat1): every thread takes ITEMS = 2 work items (same channel group g, `half` items apart) and reads each weight vector
  // from shared memory ONCE for both -- the weight reads were a third of this L1-bound kernel's data-pipe wavefronts.
  constexpr int ITEMS = (CIN == 16) ? 2 : 1;
  const int half = (ITEMS == 2) ? ((total / 2 + 3) & ~3) : total;      // multiple of 4: both items of a thread share g
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < half; t += gridDim.x * blockDim.x) {
    const int g = t & 3;
    int pixs[ITEMS];
    bool live[ITEMS];
    float acc[ITEMS][8];
    float xin[ITEMS][CIN];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int ti = t + it * half;
      live[it] = ti < total;
      pixs[it] = (live[it] ? ti : t) >> 2;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[it][c] = b_s[(c < 4 ? 0 : 12) + g * 4 + c];
      const float* ip = lat_in + (size_t)pixs[it] * CIN;
#pragma unroll
      for (int q = 0; q < CIN / 4; ++q) {
        const float4 v = ldg4(ip + 4 * q);
        xin[it][4 * q] = v.x, xin[it][4 * q + 1] = v.y, xin[it][4 * q + 2] = v.z, xin[it][4 * q + 3] = v.w;
      }
    }
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
      float4 w0, w1;
      if constexpr (CIN == 8) {   // this thread's 8 x 8 weight slice lives in registers (g is loop-invariant)
        w0 = wreg[2 * k], w1 = wreg[2 * k + 1];
      } else {
        w0 = *reinterpret_cast<const float4*>(w_s + k * 32 + g * 4);
        w1 = *reinterpret_cast<const float4*>(w_s + k * 32 + 16 + g * 4);
      }
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const float xv = xin[it][k];
        acc[it][0] = fmaf(xv, w0.x, acc[it][0]);
        acc[it][1] = fmaf(xv, w0.y, acc[it][1]);
        acc[it][2] = fmaf(xv, w0.z, acc[it][2]);
        acc[it][3] = fmaf(xv, w0.w, acc[it][3]);
        acc[it][4] = fmaf(xv, w1.x, acc[it][4]);
        acc[it][5] = fmaf(xv, w1.y, acc[it][5]);
        acc[it][6] = fmaf(xv, w1.z, acc[it][6]);
        acc[it][7] = fmaf(xv, w1.w, acc[it][7]);
      }
    }
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      if (!live[it]) continue;
      const int pix = pixs[it];
      const int x = pix % W, row = pix / W, y = row % H, n = row / H;
      // bilinear x2 with align_corners=True (ATen upsample_bilinear2d lambdas)
      const float h1r = rh * (float)y, w1r = rw * (float)x;
      const int h1 = (int)h1r, w1 = (int)w1r;
      const int h1p = (h1 < hi - 1) ? 1 : 0, w1p = (w1 < wi - 1) ? 1 : 0;
      const float h1l = h1r - (float)h1, h0l = 1.f - h1l, w1l = w1r - (float)w1, w0l = 1.f - w1l;
      const float* u00 = up_in + (((size_t)n * hi + h1) * wi + w1) * 32 + g * 4;
      const float* u01 = u00 + (size_t)w1p * 32;
      const float* u10 = u00 + (size_t)h1p * wi * 32;
      const float* u11 = u10 + (size_t)w1p * 32;
      float up[8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 a = ldg4(u00 + 16 * q), b = ldg4(u01 + 16 * q), c = ldg4(u10 + 16 * q), dd = ldg4(u11 + 16 * q);
        up[4 * q + 0] = h0l * (w0l * a.x + w1l * b.x) + h1l * (w0l * c.x + w1l * dd.x);
        up[4 * q + 1] = h0l * (w0l * a.y + w1l * b.y) + h1l * (w0l * c.y + w1l * dd.y);
        up[4 * q + 2] = h0l * (w0l * a.z + w1l * b.z) + h1l * (w0l * c.z + w1l * dd.z);
        up[4 * q + 3] = h0l * (w0l * a.w + w1l * b.w) + h1l * (w0l * c.w + w1l * dd.w);
      }
      float4* o = reinterpret_cast<float4*>(out + (size_t)pix * 32 + g * 4);
      // reference order: interpolate(x) + lateral(y)   (feature_net.py:25)
      o[0] = make_float4(up[0] + acc[it][0], up[1] + acc[it][1], up[2] + acc[it][2], up[3] + acc[it][3]);
      o[4] = make_float4(up[4] + acc[it][4], up[5] + acc[it][5], up[6] + acc[it][6], up[7] + acc[it][7]);
    }
  }
}

// cat(feat, unpreprocess(src)) -> (S,Hr,Wr,C+4); thread per output pixel
__global__ void pack_img_feat_kernel(const float* __restrict__ feat, int C, const float* __restrict__ src, int S, int H,
                                     int W, int Hr, int Wr, float* __restrict__ out) {
  const int total = S * Hr * Wr;            // 32-bit index math (launcher checks the range)
  const int CP = C + 4;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int x = t % Wr, row = t / Wr, y = row % Hr, s = row / Hr;
    const float* f = feat + (size_t)t * C;
    float* o = out + (size_t)t * CP;
    for (int q = 0; q < C / 4; ++q) reinterpret_cast<float4*>(o)[q] = ldg4(f + 4 * q);
    float rgb[3];
    for (int c = 0; c < 3; ++c) {
      const float* plane = src + ((size_t)s * 3 + c) * H * W;
      if (Hr == H && Wr == W) {
        rgb[c] = __ldg(plane + (size_t)y * W + x) * 0.5f + 0.5f;
      } else {
        // unpreprocess does the affine map first, then resizes (utils.py:609-611); the map is
        // affine so interpolating first differs only by fp32 rounding of the 4-tap blend.
        const float rh = (Hr > 1) ? (float)(H - 1) / (float)(Hr - 1) : 0.f;
        const float rw = (Wr > 1) ? (float)(W - 1) / (float)(Wr - 1) : 0.f;
        const float h1r = rh * (float)y, w1r = rw * (float)x;
        const int h1 = (int)h1r, w1 = (int)w1r;
        const int h1p = (h1 < H - 1) ? 1 : 0, w1p = (w1 < W - 1) ? 1 : 0;
        const float h1l = h1r - (float)h1, h0l = 1.f - h1l, w1l = w1r - (float)w1, w0l = 1.f - w1l;
        const float* r0 = plane + (size_t)h1 * W + w1;
        const float* r1 = r0 + (size_t)h1p * W;
        const float a = __ldg(r0) * 0.5f + 0.5f, b = __ldg(r0 + w1p) * 0.5f + 0.5f;
        const float cc = __ldg(r1) * 0.5f + 0.5f, dd = __ldg(r1 + w1p) * 0.5f + 0.5f;
        rgb[c] = h0l * (w0l * a + w1l * b) + h1l * (w0l * cc + w1l * dd);
      }
    }
    reinterpret_cast<float4*>(o + C)[0] = make_float4(rgb[0], rgb[1], rgb[2], 0.f);
  }
}

struct FeatWs {
  float *c0a, *c0, *c1a, *c1, *c2a, *c2, *f1, *f0;
  size_t bytes;
};

static FeatWs carve(void* base, int S, int H, int W) {
  FeatWs ws;
  size_t off = 0;
  auto take = [&](size_t n_floats) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align_up(n_floats * sizeof(float), 256);
    return p;
  };
  const size_t n1 = (size_t)S * H * W, n2 = n1 / 4, n4 = n1 / 16;
  ws.c0a = take(n1 * 8);
  ws.c0 = take(n1 * 8);
  ws.c1a = take(n2 * 16);
  ws.c1 = take(n2 * 16);
  ws.c2a = take(n4 * 32);
  ws.c2 = take(n4 * 32);
  ws.f1 = take(n2 * 32);
  ws.f0 = take(n1 * 32);
  ws.bytes = off;
  return ws;
}

}  // namespace enerf

extern "C" size_t enerf_feature_net_workspace_bytes(int n_views, int H, int W) {
  return enerf::carve(nullptr, n_views, H, W).bytes;
}

extern "C" int enerf_feature_net(const float* const* wts, int n_weights, const float* src_inps, int S, int H, int W,
                                 float* feat_l0, float* feat_l1, float* feat_l2, void* workspace, size_t workspace_bytes,
                                 int tensor_cores, int part, void* stream_) {
  return enerf_feature_net_packed(wts, n_weights, src_inps, S, H, W, feat_l0, feat_l1, feat_l2, nullptr, workspace, workspace_bytes, tensor_cores, part,
                                  stream_);
}

// ... and, when img_feat_rgb != NULL, the (S,H,W,12) records [level-2 features | rgb * 0.5 + 0.5 | 0] of enerf_pack_img_feat at the
// features' own resolution: written by the fused lat0 + smooth0 launch's epilogue on the tensor-core path (no extra kernel, feat_l2 is
// not re-read), by the pack kernel otherwise.  Bit-identical to enerf_pack_img_feat(feat_l2, 8, src_inps, ..., H, W).
extern "C" int enerf_feature_net_packed(const float* const* wts, int n_weights, const float* src_inps, int S, int H, int W,
                                        float* feat_l0, float* feat_l1, float* feat_l2, float* img_feat_rgb, void* workspace,
                                        size_t workspace_bytes, int tensor_cores, int part, void* stream_) {
  using namespace enerf;
  cudaStream_t stream = (cudaStream_t)stream_;
  ENERF_REQUIRE(wts && n_weights == 22, ENERF_EINVAL, "feature_net: expected 22 weight pointers, got %d", n_weights);
  ENERF_REQUIRE(src_inps && feat_l0 && feat_l1 && feat_l2 && workspace, ENERF_EINVAL, "feature_net: null pointer");
  ENERF_REQUIRE((long long)S * H * W * 4 < (1ll << 31), ENERF_EUNSUPPORTED, "feature_net: %d x %d x %d exceeds the 32-bit index range", S, H, W);
  ENERF_REQUIRE(S >= 1 && H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, ENERF_EINVAL,
                "feature_net: H=%d W=%d must be positive multiples of 4 (S=%d)", H, W, S);
  FeatWs ws = carve(workspace, S, H, W);
  ENERF_REQUIRE(workspace_bytes >= ws.bytes, ENERF_EWORKSPACE, "feature_net: workspace %zu < %zu", workspace_bytes, ws.bytes);
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  int rc;
#define RUN(call)                 \
  if ((rc = (call)) != ENERF_OK) return rc
  // stride-1 3x3 / 1x1 layers: tcgen05 implicit GEMM when tensor_cores != 0, FP32-pipe direct conv otherwise
  auto tc = [&](const char* name, int KH, int cin, int cout, int relu, const float* in, int h, int w, const float* wp,
                const float* b, float* o) {
    (void)name;
    TcConvLayer L{0, 1, KH, cin, cout, TC_PLAIN, relu};
    return tc_conv_launch(L, in, S, h, w, wp, b, nullptr, o, nullptr, cout, 0, stream);
  };
  // part 0: everything; 1: trunk (conv0.0 .. toplayer -> feat_l0); 2: pyramid tail (laterals + smooth
  // convs -> feat_l1, feat_l2; reads the trunk's c0 / c1 / feat_l0 from the same workspace).  The split
  // lets the host run the tail on a second stream next to the level-0 cost-volume chain.
  ENERF_REQUIRE(part >= 0 && part <= 2, ENERF_EINVAL, "feature_net: part %d", part);
  if (part != 2) {
  RUN(launch_conv<Conv00>("feature_net.conv0.0", src_inps, wts[0], wts[1], ws.c0a, nullptr, S, H, W, 8, 0, stream));
  if (tensor_cores) {
    RUN(tc("feature_net.conv0.1", 3, 8, 8, 1, ws.c0a, H, W, wts[2], wts[3], ws.c0));
  } else {
    RUN(launch_conv<Conv01>("feature_net.conv0.1", ws.c0a, wts[2], wts[3], ws.c0, nullptr, S, H, W, 8, 0, stream));
  }
  if (tensor_cores) {   // 5x5 stride 2: phase-tile staging (tc_conv.cuh)
    TcConvLayer L{0, 1, 5, 8, 16, TC_PLAIN, 1, 2};
    RUN(tc_conv_launch(L, ws.c0, S, H2, W2, wts[4], wts[5], nullptr, ws.c1a, nullptr, 16, 0, stream));
  } else {
    RUN(launch_conv<Conv10>("feature_net.conv1.0", ws.c0, wts[4], wts[5], ws.c1a, nullptr, S, H, W, 16, 0, stream));
  }
  if (tensor_cores) {
    RUN(tc("feature_net.conv1.1", 3, 16, 16, 1, ws.c1a, H2, W2, wts[6], wts[7], ws.c1));
  } else {
    RUN(launch_conv<Conv11>("feature_net.conv1.1", ws.c1a, wts[6], wts[7], ws.c1, nullptr, S, H2, W2, 16, 0, stream));
  }
  if (tensor_cores) {
    TcConvLayer L{0, 1, 5, 16, 32, TC_PLAIN, 1, 2};
    RUN(tc_conv_launch(L, ws.c1, S, H4, W4, wts[8], wts[9], nullptr, ws.c2a, nullptr, 32, 0, stream));
  } else {
    RUN(launch_conv<Conv20>("feature_net.conv2.0", ws.c1, wts[8], wts[9], ws.c2a, nullptr, S, H2, W2, 32, 0, stream));
  }
  if (tensor_cores) {
    RUN(tc("feature_net.conv2.1", 3, 32, 32, 1, ws.c2a, H4, W4, wts[10], wts[11], ws.c2));
    RUN(tc("feature_net.toplayer", 1, 32, 32, 0, ws.c2, H4, W4, wts[12], wts[13], feat_l0));
  } else {
    RUN(launch_conv<Conv21>("feature_net.conv2.1", ws.c2a, wts[10], wts[11], ws.c2, nullptr, S, H4, W4, 32, 0, stream));
    RUN(launch_conv<Top>("feature_net.toplayer", ws.c2, wts[12], wts[13], feat_l0, nullptr, S, H4, W4, 32, 0, stream));
  }
  }  // trunk
  if (part == 1) return ENERF_OK;
  {
    const long long total = (long long)S * H2 * W2 * 4;
    const int blocks = (int)((total / 2 + 3 + 255) / 256);      // two work items per thread (see the kernel)
    lateral_upadd_kernel<16><<<blocks, 256, 0, stream>>>(ws.c1, wts[14], wts[15], feat_l0, ws.f1, S, H2, W2);
    ENERF_CHECK_LAUNCH("feature_net.lat1");
  }
  // lat0 + smooth0: on the tensor-core path the lateral (1x1 conv + bilinear x2 + add) is computed by smooth0's producer
  // warps straight into the operand tile (tc_conv2.cu, PROD > 0): the 32-channel full-resolution map (126 MB at 512x640x3)
  // is neither written nor re-read.  Same arithmetic in the same order as lateral_upadd_kernel -> bit-identical features.
  bool fused0 = false;
  if (tensor_cores && tc_conv2_fuse_lateral()) {
    TcConvLayer L{0, 1, 3, 32, 8, TC_PLAIN, 0};
    TcLateral lat{8, ws.c0, wts[16], wts[17], ws.f1};
    if (img_feat_rgb != nullptr) lat.rgb_src = src_inps, lat.packed_out = img_feat_rgb;
    rc = tc_conv2_try_launch(L, nullptr, S, H, W, wts[20], wts[21], nullptr, feat_l2, nullptr, 8, 0, tc_fold_rule(L), stream, &lat);
    if (rc == ENERF_OK) fused0 = true;
    else if (rc != 1) return rc;
  }
  if (!fused0) {
    const long long total = (long long)S * H * W * 4;
    const int blocks = (int)((total + 1023) / 1024);     // 4 (pixel, slice) items per thread
    lateral_upadd_kernel<8><<<blocks, 256, 0, stream>>>(ws.c0, wts[16], wts[17], ws.f1, ws.f0, S, H, W);
    ENERF_CHECK_LAUNCH("feature_net.lat0");
  }
  if (tensor_cores) {
    RUN(tc("feature_net.smooth1", 3, 32, 16, 0, ws.f1, H2, W2, wts[18], wts[19], feat_l1));
    if (!fused0) RUN(tc("feature_net.smooth0", 3, 32, 8, 0, ws.f0, H, W, wts[20], wts[21], feat_l2));
  } else {
    RUN(launch_conv<Smooth1>("feature_net.smooth1", ws.f1, wts[18], wts[19], feat_l1, nullptr, S, H2, W2, 16, 0, stream));
    RUN(launch_conv<Smooth0>("feature_net.smooth0", ws.f0, wts[20], wts[21], feat_l2, nullptr, S, H, W, 8, 0, stream));
  }
  if (img_feat_rgb != nullptr && !fused0) {      // no fused launch (FP32 mode, fusion switched off, no TMA): the pack kernel
    const long long total = (long long)S * H * W;
    pack_img_feat_kernel<<<(int)((total + 255) / 256), 256, 0, stream>>>(feat_l2, 8, src_inps, S, H, W, H, W, img_feat_rgb);
    ENERF_CHECK_LAUNCH("feature_net.pack_img_feat");
  }
#undef RUN
  return ENERF_OK;
}

extern "C" int enerf_pack_img_feat(const float* feat, int C, const float* src_inps, int S, int H, int W, int Hr, int Wr,
                                   float* out, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(feat && src_inps && out, ENERF_EINVAL, "pack_img_feat: null pointer");
  ENERF_REQUIRE(C % 4 == 0 && C > 0, ENERF_EINVAL, "pack_img_feat: C=%d must be a multiple of 4", C);
  ENERF_REQUIRE((long long)S * Hr * Wr < (1ll << 31), ENERF_EUNSUPPORTED, "pack_img_feat: frame too large for 32-bit indices");
  const long long total = (long long)S * Hr * Wr;
  const int blocks = (int)((total + 255) / 256);
  pack_img_feat_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(feat, C, src_inps, S, H, W, Hr, Wr, out);
  ENERF_CHECK_LAUNCH("pack_img_feat");
  return ENERF_OK;
}
