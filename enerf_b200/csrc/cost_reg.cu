// cost_reg.cu -- MinCostRegNet.forward (/root/reference/lib/networks/enerf/cost_reg_net.py:75-86)
// and CostRegNet.forward (:35-48): 3-D U-Net over the variance volume, channels-last (D,h,w,C).
//
//  * full-resolution layers (conv0, the fused feat_conv+depth_conv head): tiled direct conv
//    (conv.cuh), one CTA per 4x8x32 voxel tile, FP32-pipe bound;
//  * coarse layers (1/2 .. 1/8 resolution, 1k-80k voxels): `conv3d_small` -- a thread owns one
//    voxel x 8 output channels x one 8-channel slice of Cin; the Cin slices of a voxel sit in
//    adjacent lanes and are combined with warp shuffles, so even the (1,32,40) x 64-channel bottom
//    of the L1 net spreads over >2500 warps;
//  * ConvTranspose3d(k3,s2,p1,op1)+BN+skip (cost_reg_net.py:19-33,43-46): 8-phase sub-pixel form,
//    a thread owns a 2x2x2 output block (all 8 parities => exactly the 27 (tap,input) pairs of
//    the 2x2x2 input neighbourhood), no divergence, skip-add fused.
#include "conv.cuh"
#include "tc_conv.cuh"

namespace enerf {

//                         CIN COUT KD KH ST TZ TY TX  PY PX COT CCH RELU  PLANAR HEAD
using C0_32 = ConvTraits<32, 8, 3, 3, 1, 4, 8, 32, 2, 4, 8, 4, true, false>;
using C0_16 = ConvTraits<16, 8, 3, 3, 1, 4, 8, 32, 2, 4, 8, 4, true, false>;
using C0_8 = ConvTraits<8, 8, 3, 3, 1, 4, 8, 32, 2, 4, 8, 4, true, false>;
using Head9 = ConvTraits<8, 9, 3, 3, 1, 4, 8, 32, 2, 4, 9, 4, false, false, true>;
using Head1 = ConvTraits<8, 1, 3, 3, 1, 4, 8, 32, 2, 4, 1, 4, false, false>;

// ---- coarse-layer forward conv (3x3x3, pad 1, stride 1|2), relu(conv + bias) -------------------
// thread = (voxel, 8-channel output group g, 8-channel input slice k); lanes: k fastest.
template <int CIN, int COUT, int STRIDE>
__global__ void __launch_bounds__(256) conv3d_small_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           int Di, int Hi, int Wi, int Do, int Ho, int Wo) {
  constexpr int KS = CIN / 8, NG = COUT / 8;
  const long long total = (long long)Do * Ho * Wo * NG * KS;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < total;
  const long long tt = active ? t : 0;
  const int k = (int)(tt % KS);
  const int g = (int)((tt / KS) % NG);
  const long long vox = tt / (KS * NG);
  const int ox = (int)(vox % Wo), oy = (int)((vox / Wo) % Ho), oz = (int)(vox / ((long long)Wo * Ho));
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll 1
  for (int kz = 0; kz < 3; ++kz) {
    const int iz = oz * STRIDE - 1 + kz;
    if (iz < 0 || iz >= Di) continue;
#pragma unroll 1
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * STRIDE - 1 + ky;
      if (iy < 0 || iy >= Hi) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * STRIDE - 1 + kx;
        if (ix < 0 || ix >= Wi) continue;
        const float* ip = in + (((size_t)iz * Hi + iy) * Wi + ix) * CIN + k * 8;
        const float4 a = ldg4(ip), b = ldg4(ip + 4);
        const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const float* wp = w + ((size_t)((kz * 3 + ky) * 3 + kx) * CIN + k * 8) * COUT + g * 8;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const float4 w0 = ldg4(wp + (size_t)ci * COUT), w1 = ldg4(wp + (size_t)ci * COUT + 4);
          acc[0] = fmaf(xv[ci], w0.x, acc[0]);
          acc[1] = fmaf(xv[ci], w0.y, acc[1]);
          acc[2] = fmaf(xv[ci], w0.z, acc[2]);
          acc[3] = fmaf(xv[ci], w0.w, acc[3]);
          acc[4] = fmaf(xv[ci], w1.x, acc[4]);
          acc[5] = fmaf(xv[ci], w1.y, acc[5]);
          acc[6] = fmaf(xv[ci], w1.z, acc[6]);
          acc[7] = fmaf(xv[ci], w1.w, acc[7]);
        }
      }
    }
  }
  // combine the KS input slices (adjacent lanes; total is a multiple of KS so groups never straddle)
#pragma unroll
  for (int off = 1; off < KS; off <<= 1)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
  if (active && k == 0) {
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = fmaxf(acc[c] + __ldg(bias + g * 8 + c), 0.f);
    float4* o = reinterpret_cast<float4*>(out + (size_t)vox * COUT + g * 8);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ---- ConvTranspose3d(k=3, s=2, p=1, output_padding=1) + folded BN + skip add ---------------------
// out[o] = sum_{i,k : o = 2i - 1 + k} in[i] w[k];  o = 2b + e:  e=0 -> (k=1,i=b);  e=1 -> (k=0,i=b+1),(k=2,i=b)
// thread = (2x2x2 output block b, 8-channel output group g, 8-channel input slice k)
template <int CIN, int COUT>
__global__ void __launch_bounds__(128) deconv3d_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* __restrict__ skip,
                                                       float* __restrict__ out, int Di, int Hi, int Wi) {
  constexpr int KS = CIN / 8, NG = COUT / 8;
  const long long total = (long long)Di * Hi * Wi * NG * KS;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < total;
  const long long tt = active ? t : 0;
  const int k = (int)(tt % KS);
  const int g = (int)((tt / KS) % NG);
  const long long blk = tt / (KS * NG);
  const int bx = (int)(blk % Wi), by = (int)((blk / Wi) % Hi), bz = (int)(blk / ((long long)Wi * Hi));
  float acc[8][8];  // [output parity ez*4+ey*2+ex][co]
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[p][c] = 0.f;

#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iz = bz + dz, iy = by + dy, ix = bx + dx;
        if (iz >= Di || iy >= Hi || ix >= Wi) continue;  // warp-divergent only at the far faces
        const float* ip = in + (((size_t)iz * Hi + iy) * Wi + ix) * CIN + k * 8;
        const float4 a = ldg4(ip), b = ldg4(ip + 4);
        const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        // input offset d in {0,1} feeds output parity e with tap: d=0: (e=0,k=1),(e=1,k=2); d=1: (e=1,k=0)
#pragma unroll
        for (int ez = dz; ez < 2; ++ez)
#pragma unroll
          for (int ey = dy; ey < 2; ++ey)
#pragma unroll
            for (int ex = dx; ex < 2; ++ex) {
              const int kz = dz ? 0 : (ez ? 2 : 1), ky = dy ? 0 : (ey ? 2 : 1), kx = dx ? 0 : (ex ? 2 : 1);
              const float* wp = w + ((size_t)((kz * 3 + ky) * 3 + kx) * CIN + k * 8) * COUT + g * 8;
              const int p = ez * 4 + ey * 2 + ex;
#pragma unroll
              for (int ci = 0; ci < 8; ++ci) {
                const float4 w0 = ldg4(wp + (size_t)ci * COUT), w1 = ldg4(wp + (size_t)ci * COUT + 4);
                acc[p][0] = fmaf(xv[ci], w0.x, acc[p][0]);
                acc[p][1] = fmaf(xv[ci], w0.y, acc[p][1]);
                acc[p][2] = fmaf(xv[ci], w0.z, acc[p][2]);
                acc[p][3] = fmaf(xv[ci], w0.w, acc[p][3]);
                acc[p][4] = fmaf(xv[ci], w1.x, acc[p][4]);
                acc[p][5] = fmaf(xv[ci], w1.y, acc[p][5]);
                acc[p][6] = fmaf(xv[ci], w1.z, acc[p][6]);
                acc[p][7] = fmaf(xv[ci], w1.w, acc[p][7]);
              }
            }
      }
#pragma unroll
  for (int off = 1; off < KS; off <<= 1)
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[p][c] += __shfl_xor_sync(0xffffffffu, acc[p][c], off);
  if (active && k == 0) {
    const int Ho = 2 * Hi, Wo = 2 * Wi;
    float bv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bv[c] = __ldg(bias + g * 8 + c);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int oz = 2 * bz + (p >> 2), oy = 2 * by + ((p >> 1) & 1), ox = 2 * bx + (p & 1);
      const size_t o = (((size_t)oz * Ho + oy) * Wo + ox) * COUT + g * 8;
      const float4 s0 = ldg4(skip + o), s1 = ldg4(skip + o + 4);
      float4* op = reinterpret_cast<float4*>(out + o);
      // reference: skip + BN(deconv(x))   (cost_reg_net.py:41-46)
      op[0] = make_float4(s0.x + (acc[p][0] + bv[0]), s0.y + (acc[p][1] + bv[1]), s0.z + (acc[p][2] + bv[2]),
                          s0.w + (acc[p][3] + bv[3]));
      op[1] = make_float4(s1.x + (acc[p][4] + bv[4]), s1.y + (acc[p][5] + bv[5]), s1.z + (acc[p][6] + bv[6]),
                          s1.w + (acc[p][7] + bv[7]));
    }
  }
}

template <int CIN, int COUT, int STRIDE>
static int launch_small(const char* name, const float* in, const float* w, const float* b, float* out, int Di, int Hi,
                        int Wi, cudaStream_t stream) {
  const int Do = (Di + 2 - 3) / STRIDE + 1, Ho = (Hi + 2 - 3) / STRIDE + 1, Wo = (Wi + 2 - 3) / STRIDE + 1;
  const long long total = (long long)Do * Ho * Wo * (COUT / 8) * (CIN / 8);
  conv3d_small_kernel<CIN, COUT, STRIDE><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, w, b, out, Di, Hi, Wi,
                                                                                             Do, Ho, Wo);
  ENERF_CHECK_LAUNCH(name);
  return ENERF_OK;
}

template <int CIN, int COUT>
static int launch_deconv(const char* name, const float* in, const float* w, const float* b, const float* skip, float* out,
                         int Di, int Hi, int Wi, cudaStream_t stream) {
  const long long total = (long long)Di * Hi * Wi * (COUT / 8) * (CIN / 8);
  deconv3d_kernel<CIN, COUT><<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(in, w, b, skip, out, Di, Hi, Wi);
  ENERF_CHECK_LAUNCH(name);
  return ENERF_OK;
}

struct RegWs {
  float *c0, *c1, *c2, *c3, *c4, *c5, *c6, *y4, *y2, *y0;
  size_t bytes;
};

static RegWs carve_reg(void* base, int deep, int D, int h, int w) {
  RegWs ws;
  size_t off = 0;
  auto take = [&](size_t n_floats) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align_up(n_floats * sizeof(float), 256);
    return p;
  };
  const size_t v1 = (size_t)D * h * w, v2 = v1 / 8, v4 = v2 / 8, v8 = v4 / 8;
  ws.c0 = take(v1 * 8);
  ws.c1 = take(v2 * 16);
  ws.c2 = take(v2 * 16);
  ws.c3 = take(v4 * 32);
  ws.c4 = take(v4 * 32);
  ws.c5 = deep ? take(v8 * 64) : nullptr;
  ws.c6 = deep ? take(v8 * 64) : nullptr;
  ws.y4 = deep ? take(v4 * 32) : nullptr;
  ws.y2 = take(v2 * 16);
  ws.y0 = take(v1 * 8);
  ws.bytes = off;
  return ws;
}

}  // namespace enerf

extern "C" size_t enerf_cost_reg_workspace_bytes(int deep, int D, int h, int w) {
  return enerf::carve_reg(nullptr, deep, D, h, w).bytes;
}

extern "C" int enerf_cost_reg(const float* const* wts, int n_weights, int deep, int in_ch, const float* variance, int D,
                              int h, int w, float* feat_vol, float* depth_prob, void* workspace, size_t workspace_bytes,
                              int tensor_cores, void* stream_) {
  using namespace enerf;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int expect = deep ? 21 : 15;
  ENERF_REQUIRE(wts && n_weights == expect, ENERF_EINVAL, "cost_reg: expected %d weight pointers, got %d", expect, n_weights);
  ENERF_REQUIRE(variance && depth_prob && workspace, ENERF_EINVAL, "cost_reg: null pointer");
  ENERF_REQUIRE(in_ch == 8 || in_ch == 16 || in_ch == 32, ENERF_EUNSUPPORTED, "cost_reg: in_ch %d not in {8,16,32}", in_ch);
  const int div = deep ? 8 : 4;
  ENERF_REQUIRE(D > 0 && h > 0 && w > 0 && D % div == 0 && h % div == 0 && w % div == 0, ENERF_EINVAL,
                "cost_reg: volume %dx%dx%d must be divisible by %d (skip connections, cost_reg_net.py:%s)", D, h, w, div,
                deep ? "41-46" : "80-83");
  RegWs ws = carve_reg(workspace, deep, D, h, w);
  ENERF_REQUIRE(workspace_bytes >= ws.bytes, ENERF_EWORKSPACE, "cost_reg: workspace %zu < %zu", workspace_bytes, ws.bytes);
  int rc;
#define RUN(call)                 \
  if ((rc = (call)) != ENERF_OK) return rc
  const int D2 = D / 2, h2 = h / 2, w2 = w / 2, D4 = D / 4, h4 = h / 4, w4 = w / 4, D8 = D / 8, h8 = h / 8, w8 = w / 8;
  // tcgen05 path for every stride-1 and transposed layer (stride-2 layers stay on the FP32 pipe)
  auto tc_cbr = [&](int cin, int cout, const float* in, int d_, int h_, int w_, const float* wp, const float* b, float* o) {
    TcConvLayer L{0, 3, 3, cin, cout, TC_PLAIN, 1};
    return tc_conv_launch(L, in, d_, h_, w_, wp, b, nullptr, o, nullptr, cout, 0, stream);
  };
  auto tc_dec = [&](int cin, int cout, const float* in, int d_, int h_, int w_, const float* wp, const float* b, const float* skip,
                    float* o) {
    TcConvLayer L{1, 3, 3, cin, cout, TC_DECONV, 0};
    return tc_conv_launch(L, in, d_, h_, w_, wp, b, skip, o, nullptr, cout, 0, stream);
  };
  if (tensor_cores) {
    RUN(tc_cbr(in_ch, 8, variance, D, h, w, wts[0], wts[1], ws.c0));
  } else if (in_ch == 32) {
    RUN(launch_conv<C0_32>("cost_reg.conv0", variance, wts[0], wts[1], ws.c0, nullptr, D, h, w, 8, 0, stream));
  } else if (in_ch == 16) {
    RUN(launch_conv<C0_16>("cost_reg.conv0", variance, wts[0], wts[1], ws.c0, nullptr, D, h, w, 8, 0, stream));
  } else {
    RUN(launch_conv<C0_8>("cost_reg.conv0", variance, wts[0], wts[1], ws.c0, nullptr, D, h, w, 8, 0, stream));
  }
  auto tc_s2 = [&](int cin, int cout, const float* in, int do_, int ho_, int wo_, const float* wp, const float* b, float* o) {
    TcConvLayer L{0, 3, 3, cin, cout, TC_PLAIN, 1, 2};        // (do_,ho_,wo_) = OUTPUT grid
    return tc_conv_launch(L, in, do_, ho_, wo_, wp, b, nullptr, o, nullptr, cout, 0, stream);
  };
  if (tensor_cores) {
    RUN(tc_s2(8, 16, ws.c0, D2, h2, w2, wts[2], wts[3], ws.c1));
  } else {
    RUN((launch_small<8, 16, 2>("cost_reg.conv1", ws.c0, wts[2], wts[3], ws.c1, D, h, w, stream)));
  }
  if (tensor_cores) {
    RUN(tc_cbr(16, 16, ws.c1, D2, h2, w2, wts[4], wts[5], ws.c2));
  } else {
    RUN((launch_small<16, 16, 1>("cost_reg.conv2", ws.c1, wts[4], wts[5], ws.c2, D2, h2, w2, stream)));
  }
  if (tensor_cores) {
    RUN(tc_s2(16, 32, ws.c2, D4, h4, w4, wts[6], wts[7], ws.c3));
  } else {
    RUN((launch_small<16, 32, 2>("cost_reg.conv3", ws.c2, wts[6], wts[7], ws.c3, D2, h2, w2, stream)));
  }
  if (tensor_cores) {
    RUN(tc_cbr(32, 32, ws.c3, D4, h4, w4, wts[8], wts[9], ws.c4));
  } else {
    RUN((launch_small<32, 32, 1>("cost_reg.conv4", ws.c3, wts[8], wts[9], ws.c4, D4, h4, w4, stream)));
  }
  const float* x4 = ws.c4;
  int wi = 10;
  if (deep) {
    if (tensor_cores) {
      RUN(tc_s2(32, 64, ws.c4, D8, h8, w8, wts[10], wts[11], ws.c5));
      RUN(tc_cbr(64, 64, ws.c5, D8, h8, w8, wts[12], wts[13], ws.c6));
      RUN(tc_dec(64, 32, ws.c6, D8, h8, w8, wts[14], wts[15], ws.c4, ws.y4));
    } else {
      RUN((launch_small<32, 64, 2>("cost_reg.conv5", ws.c4, wts[10], wts[11], ws.c5, D4, h4, w4, stream)));
      RUN((launch_small<64, 64, 1>("cost_reg.conv6", ws.c5, wts[12], wts[13], ws.c6, D8, h8, w8, stream)));
      RUN((launch_deconv<64, 32>("cost_reg.conv7", ws.c6, wts[14], wts[15], ws.c4, ws.y4, D8, h8, w8, stream)));
    }
    x4 = ws.y4;
    wi = 16;
  }
  if (tensor_cores) {
    RUN(tc_dec(32, 16, x4, D4, h4, w4, wts[wi], wts[wi + 1], ws.c2, ws.y2));
    RUN(tc_dec(16, 8, ws.y2, D2, h2, w2, wts[wi + 2], wts[wi + 3], ws.c0, ws.y0));
  } else {
    RUN((launch_deconv<32, 16>("cost_reg.conv9", x4, wts[wi], wts[wi + 1], ws.c2, ws.y2, D4, h4, w4, stream)));
    RUN((launch_deconv<16, 8>("cost_reg.conv11", ws.y2, wts[wi + 2], wts[wi + 3], ws.c0, ws.y0, D2, h2, w2, stream)));
  }
  if (tensor_cores) {
    TcConvLayer Lh{0, 3, 3, 8, feat_vol ? 9 : 1, feat_vol ? TC_HEAD : TC_SINGLE, 0};
    RUN(tc_conv_launch(Lh, ws.y0, D, h, w, wts[wi + 4], nullptr, nullptr, feat_vol ? feat_vol : depth_prob, depth_prob, 8, 0, stream));
  } else {
    if (feat_vol) {
      RUN(launch_conv<Head9>("cost_reg.head9", ws.y0, wts[wi + 4], nullptr, feat_vol, depth_prob, D, h, w, 8, 0, stream));
    } else {
      // depth_conv only: the caller passes the head as a [27][8][1] tensor (see enerf_b200.h)
      RUN(launch_conv<Head1>("cost_reg.head1", ws.y0, wts[wi + 4], nullptr, depth_prob, nullptr, D, h, w, 1, 0, stream));
    }
  }
#undef RUN
  return ENERF_OK;
}
