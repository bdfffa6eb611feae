// conv.cuh -- direct convolution on the FP32 pipe for the small-channel CNNs of ENeRF.
//
// One template covers every forward convolution of FeatureNet (feature_net.py:7-22, 2-D: KD=1 and
// the S source images mapped to the depth axis) and (Min)CostRegNet (cost_reg_net.py:7-33,
// 3x3x3, stride 1/2).  Channels-last activations, BN folded into {w,bias} on the host.
//
// Mapping: a CTA owns a TZ x TY x TX output tile for ALL output channels.  The input halo tile is
// staged in shared memory CCHUNK input channels at a time, transposed to channel-planar so that a
// thread's x-strip is contiguous (LDS.128, conflict-free per quarter warp); the weights of the
// chunk are staged as [tap][ci][cout] and read as warp-broadcast LDS.128.  A thread accumulates a
// PY x PX pixel patch x CO_T channels in registers, so one weight vector feeds PY*PX FMAs per
// lane and one input strip feeds KH*CO_T: the FMA:LDS ratio is ~20:1 and the kernel is bound by
// the FP32 pipe, which is the right roofline for Cout in {8..64} (SURVEY.md section 7 "tiny channel counts").
#pragma once
#include "common.cuh"

namespace enerf {

struct ConvDims {
  int Di, Hi, Wi;     // input extent
  int Do, Ho, Wo;     // output extent
  int out_cstride;    // channel stride of the output tensor (>= COUT)
  int out_coff;       // first output channel
};

// Epilogues ---------------------------------------------------------------------------------------
struct EpiBias {         // y = conv + bias
  static constexpr bool kRelu = false;
};
struct EpiBiasRelu {     // y = relu(conv + bias)
  static constexpr bool kRelu = true;
};

template <int CIN_, int COUT_, int KD_, int KH_, int STRIDE_, int TZ_, int TY_, int TX_, int PY_, int PX_, int CO_T_,
          int CCHUNK_, bool RELU_, bool IN_PLANAR_, bool HEAD_ = false>
struct ConvTraits {
  static constexpr int CIN = CIN_, COUT = COUT_, KD = KD_, KH = KH_, STRIDE = STRIDE_;
  static constexpr int TZ = TZ_, TY = TY_, TX = TX_, PY = PY_, PX = PX_, CO_T = CO_T_, CCHUNK = CCHUNK_;
  static constexpr bool RELU = RELU_, IN_PLANAR = IN_PLANAR_, HEAD = HEAD_;
  static constexpr int SZ = (KD == 1) ? 1 : STRIDE;          // 2-D convs do not stride over images
  static constexpr int PADZ = KD / 2, PAD = KH / 2;
  static constexpr int NX = TX / PX, NY = TY / PY, NG = COUT / CO_T;
  static constexpr int THREADS = NX * NY * TZ * NG;
  static constexpr int IZ = (TZ - 1) * SZ + KD, IY = (TY - 1) * STRIDE + KH, IX = (TX - 1) * STRIDE + KH;
  static constexpr int LEN = (PX - 1) * STRIDE + KH;          // input strip a thread needs per row
  static constexpr int LEN4 = (LEN + 3) / 4;
  static constexpr int IXP = (((TX - PX) * STRIDE + LEN4 * 4) + 3) / 4 * 4;
  static constexpr int ROWS = (PY - 1) * STRIDE + KH;
  static constexpr int PLANE = IZ * IY * IXP;
  static constexpr int TAPS = KD * KH * KH;
  static constexpr int CO_PAD = (COUT + 3) / 4 * 4;           // smem weight row (floats)
  static constexpr int COT_PAD = (CO_T + 3) / 4 * 4;
  static constexpr int NCHUNK = CIN / CCHUNK;
  static constexpr size_t SMEM = (size_t)(CCHUNK * PLANE + TAPS * CCHUNK * CO_PAD) * sizeof(float);
  static_assert(CIN % CCHUNK == 0, "CIN must be a multiple of CCHUNK");
  static_assert(COUT % CO_T == 0, "COUT must be a multiple of CO_T");
  static_assert(TX % PX == 0 && TY % PY == 0, "tile / patch mismatch");
  static_assert(IN_PLANAR || CCHUNK % 4 == 0, "channels-last input is read as float4");
  static_assert((PX * STRIDE) % 4 == 0, "strip start must be 16-byte aligned");
};

// in : channels-last (Di,Hi,Wi,CIN)   [IN_PLANAR: (Di,CIN,Hi,Wi)]
// w  : [TAPS][CIN][COUT]   bias: [COUT] or nullptr
// out: channels-last (Do,Ho,Wo,out_cstride), channels [out_coff, out_coff+COUT)
//      HEAD: COUT==9 -> channels 0..7 to `out` (stride 8), channel 8 to `out2` (Do,Ho,Wo)
// skip: optional tensor added after bias (same layout as out), nullptr if none
template <class T>
__global__ void __launch_bounds__(T::THREADS) conv_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ out,
                                                          float* __restrict__ out2, ConvDims d) {
  extern __shared__ __align__(16) float smem[];
  float* in_s = smem;                          // [CCHUNK][IZ][IY][IXP]
  float* w_s = smem + T::CCHUNK * T::PLANE;    // [TAPS][CCHUNK][CO_PAD]

  const int tid = threadIdx.x;
  const int tx = tid % T::NX;
  const int ty = (tid / T::NX) % T::NY;
  const int tz = (tid / (T::NX * T::NY)) % T::TZ;
  const int tg = tid / (T::NX * T::NY * T::TZ);

  const int ox0 = blockIdx.x * T::TX, oy0 = blockIdx.y * T::TY, oz0 = blockIdx.z * T::TZ;
  const int ix0 = ox0 * T::STRIDE - T::PAD, iy0 = oy0 * T::STRIDE - T::PAD, iz0 = oz0 * T::SZ - T::PADZ;

  float acc[T::PY][T::PX][T::CO_T];
#pragma unroll
  for (int a = 0; a < T::PY; ++a)
#pragma unroll
    for (int b = 0; b < T::PX; ++b)
#pragma unroll
      for (int c = 0; c < T::CO_T; ++c) acc[a][b][c] = 0.f;

  for (int chunk = 0; chunk < T::NCHUNK; ++chunk) {
    const int c0 = chunk * T::CCHUNK;
    if (chunk > 0) __syncthreads();
    // ---- stage the input halo tile (zero padded) ----
    if constexpr (T::IN_PLANAR) {
      constexpr int N = T::CCHUNK * T::IZ * T::IY * T::IX;
      for (int e = tid; e < N; e += T::THREADS) {
        const int x = e % T::IX, y = (e / T::IX) % T::IY, z = (e / (T::IX * T::IY)) % T::IZ, c = e / (T::IX * T::IY * T::IZ);
        const int gx = ix0 + x, gy = iy0 + y, gz = iz0 + z;
        float v = 0.f;
        if (gx >= 0 && gx < d.Wi && gy >= 0 && gy < d.Hi && gz >= 0 && gz < d.Di)
          v = __ldg(in + (((size_t)gz * T::CIN + c0 + c) * d.Hi + gy) * d.Wi + gx);
        in_s[c * T::PLANE + (z * T::IY + y) * T::IXP + x] = v;
      }
    } else {
      constexpr int Q = T::CCHUNK / 4;
      constexpr int N = Q * T::IZ * T::IY * T::IX;
      for (int e = tid; e < N; e += T::THREADS) {
        const int x = e % T::IX, y = (e / T::IX) % T::IY, z = (e / (T::IX * T::IY)) % T::IZ, q = e / (T::IX * T::IY * T::IZ);
        const int gx = ix0 + x, gy = iy0 + y, gz = iz0 + z;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx >= 0 && gx < d.Wi && gy >= 0 && gy < d.Hi && gz >= 0 && gz < d.Di)
          v = ldg4(in + (((size_t)gz * d.Hi + gy) * d.Wi + gx) * T::CIN + c0 + q * 4);
        float* p = in_s + (q * 4) * T::PLANE + (z * T::IY + y) * T::IXP + x;
        p[0] = v.x;
        p[T::PLANE] = v.y;
        p[2 * T::PLANE] = v.z;
        p[3 * T::PLANE] = v.w;
      }
    }
    // ---- stage this chunk's weights: w_s[tap][ci][co] <- w[tap][c0+ci][co] ----
    {
      constexpr int N = T::TAPS * T::CCHUNK * T::CO_PAD;
      for (int e = tid; e < N; e += T::THREADS) {
        const int co = e % T::CO_PAD, ci = (e / T::CO_PAD) % T::CCHUNK, tap = e / (T::CO_PAD * T::CCHUNK);
        w_s[e] = (co < T::COUT) ? __ldg(w + ((size_t)tap * T::CIN + c0 + ci) * T::COUT + co) : 0.f;
      }
    }
    __syncthreads();

    // ---- accumulate ----
    const float* in_t = in_s + ((tz * T::SZ) * T::IY + ty * T::PY * T::STRIDE) * T::IXP + tx * T::PX * T::STRIDE;
    const float* w_t = w_s + tg * T::CO_T;
#pragma unroll 1
    for (int ci = 0; ci < T::CCHUNK; ++ci) {
#pragma unroll
      for (int kd = 0; kd < T::KD; ++kd) {
        float xin[T::ROWS][T::LEN4 * 4];
#pragma unroll
        for (int r = 0; r < T::ROWS; ++r) {
          const float4* row = reinterpret_cast<const float4*>(in_t + ci * T::PLANE + (kd * T::IY + r) * T::IXP);
#pragma unroll
          for (int q = 0; q < T::LEN4; ++q) {
            const float4 v = row[q];
            xin[r][4 * q + 0] = v.x;
            xin[r][4 * q + 1] = v.y;
            xin[r][4 * q + 2] = v.z;
            xin[r][4 * q + 3] = v.w;
          }
        }
#pragma unroll
        for (int ky = 0; ky < T::KH; ++ky) {
#pragma unroll
          for (int kx = 0; kx < T::KH; ++kx) {
            float wv[T::COT_PAD];
            const float* wp = w_t + (((kd * T::KH + ky) * T::KH + kx) * T::CCHUNK + ci) * T::CO_PAD;
            if constexpr (T::CO_T % 4 == 0 || T::NG == 1) {
#pragma unroll
              for (int q = 0; q < T::COT_PAD / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(wp + 4 * q);
                wv[4 * q + 0] = v.x;
                wv[4 * q + 1] = v.y;
                wv[4 * q + 2] = v.z;
                wv[4 * q + 3] = v.w;
              }
            } else {
#pragma unroll
              for (int c = 0; c < T::CO_T; ++c) wv[c] = wp[c];
            }
#pragma unroll
            for (int a = 0; a < T::PY; ++a)
#pragma unroll
              for (int b = 0; b < T::PX; ++b) {
                const float xv = xin[a * T::STRIDE + ky][b * T::STRIDE + kx];
#pragma unroll
                for (int c = 0; c < T::CO_T; ++c) acc[a][b][c] = fmaf(xv, wv[c], acc[a][b][c]);
              }
          }
        }
      }
    }
  }

  // ---- epilogue ----
  const int oz = oz0 + tz;
  if (oz >= d.Do) return;
  float bv[T::CO_T];
#pragma unroll
  for (int c = 0; c < T::CO_T; ++c) bv[c] = bias ? __ldg(bias + tg * T::CO_T + c) : 0.f;
#pragma unroll
  for (int a = 0; a < T::PY; ++a) {
    const int oy = oy0 + ty * T::PY + a;
    if (oy >= d.Ho) continue;
#pragma unroll
    for (int b = 0; b < T::PX; ++b) {
      const int ox = ox0 + tx * T::PX + b;
      if (ox >= d.Wo) continue;
      const size_t pix = ((size_t)oz * d.Ho + oy) * d.Wo + ox;
      float v[T::CO_T];
#pragma unroll
      for (int c = 0; c < T::CO_T; ++c) {
        v[c] = acc[a][b][c] + bv[c];
        if (T::RELU) v[c] = fmaxf(v[c], 0.f);
      }
      if constexpr (T::HEAD) {
        static_assert(!T::HEAD || (T::COUT == 9 && T::CO_T == 9), "head = feat_conv(8) + depth_conv(1)");
        float4* o = reinterpret_cast<float4*>(out + pix * 8);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
        out2[pix] = v[8];
      } else if constexpr (T::CO_T % 4 == 0) {
        float4* o = reinterpret_cast<float4*>(out + pix * d.out_cstride + d.out_coff + tg * T::CO_T);
#pragma unroll
        for (int q = 0; q < T::CO_T / 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else {
#pragma unroll
        for (int c = 0; c < T::CO_T; ++c) out[pix * d.out_cstride + d.out_coff + tg * T::CO_T + c] = v[c];
      }
    }
  }
}

template <class T>
int launch_conv(const char* name, const float* in, const float* w, const float* bias, float* out, float* out2,
                int Di, int Hi, int Wi, int out_cstride, int out_coff, cudaStream_t stream) {
  ConvDims d;
  d.Di = Di, d.Hi = Hi, d.Wi = Wi;
  d.Do = (T::KD == 1) ? Di : (Di + 2 * T::PADZ - T::KD) / T::SZ + 1;
  d.Ho = (Hi + 2 * T::PAD - T::KH) / T::STRIDE + 1;
  d.Wo = (Wi + 2 * T::PAD - T::KH) / T::STRIDE + 1;
  d.out_cstride = out_cstride, d.out_coff = out_coff;
  static PerDeviceSize attr_set;   // the attribute is per device
  if (attr_set.cur() < T::SMEM) {
    cudaError_t e = cudaFuncSetAttribute(conv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute(%zu B smem): %s", name, T::SMEM, cudaGetErrorString(e));
      return ENERF_ECUDA;
    }
    attr_set.cur() = T::SMEM;
  }
  dim3 grid(ceil_div(d.Wo, T::TX), ceil_div(d.Ho, T::TY), ceil_div(d.Do, T::TZ));
  conv_kernel<T><<<grid, T::THREADS, T::SMEM, stream>>>(in, w, bias, out, out2, d);
  ENERF_CHECK_LAUNCH(name);
  return ENERF_OK;
}

}  // namespace enerf
