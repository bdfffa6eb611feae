// tc_conv.cu -- host side of the tcgen05 implicit-GEMM convolution (kernel + design: tc_conv.cuh):
// tile selection, TMA tensor-map encoding (cuTensorMapEncodeTiled through the runtime's driver
// entry point -- no link-time dependency on libcuda), launch, and the C-ABI test/diagnostic entry.
#include <cudaTypedefs.h>

#include "tc_conv.cuh"

namespace enerf {

struct TcConvParams {
  int Dn, Hn, Wn;        // row grid (output grid of a conv; INPUT grid of a transposed conv)
  int TZ, TY, TX;        // tile
  int IZ, IY, IX;        // tile + halo
  int oz, oy, ox;        // halo origin = tile origin - (oz,oy,ox)
  int n_taps;
  int tap_off[27];       // linear offsets (pixels) inside the halo tile
  int n_stages;          // Cin / 8
  int N;                 // MMA N (multiple of 16, <= 256)
  int n_mt;              // 128-row M-tiles per CTA
  int cout;              // real channels (per parity for TC_DECONV)
  int relu, mode;
  int out_cstride, out_coff;
  uint32_t tmem_cols;
  const float* wpack;    // [stage][tap][2][N][4] TF32
  const float* bias;     // [cout] or nullptr
  const float* skip;     // TC_DECONV: tensor added to the result (same layout as out)
  float* out;
  float* out2;           // TC_HEAD: depth_prob
};

__global__ void __launch_bounds__(128) tc_conv_kernel(const __grid_constant__ CUtensorMap tmap, const TcConvParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  const int npix = P.IZ * P.IY * P.IX;
  const uint32_t a_bytes = (uint32_t)npix * 32u;                   // two 4-channel planes
  const uint32_t w_bytes = (uint32_t)P.n_taps * (uint32_t)P.N * 32u;
  const uint32_t stage_bytes = (a_bytes + w_bytes + 127u) & ~127u;
  unsigned char* stage0 = smem_raw;

  if (t == 0) {
    tc::prefetch_tmap(&tmap);
    tc::mbar_init(&full_bar[0], 1);
    tc::mbar_init(&full_bar[1], 1);
    tc::mbar_init(&empty_bar[0], 1);
    tc::mbar_init(&empty_bar[1], 1);
    tc::mbar_init(&done_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, P.tmem_cols);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  const int x0 = blockIdx.x * P.TX, y0 = blockIdx.y * P.TY, z0 = blockIdx.z * P.TZ;

  if (t == 32) {
    // ---------------- TMA producer ----------------
    for (int st = 0; st < P.n_stages; ++st) {
      const int slot = st & 1;
      if (st >= 2) tc::mbar_wait(&empty_bar[slot], (uint32_t)(((st - 2) >> 1) & 1));
      unsigned char* sa = stage0 + (size_t)slot * stage_bytes;
      tc::mbar_expect_tx(&full_bar[slot], a_bytes + w_bytes);
      tc::tma_load_5d(sa, &tmap, 0, x0 - P.ox, y0 - P.oy, z0 - P.oz, 2 * st, &full_bar[slot]);
      tc::tma_load_1d(sa + a_bytes, P.wpack + (size_t)st * (w_bytes / 4), w_bytes, &full_bar[slot]);
    }
  } else if (t == 0) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = tc::idesc_tf32(128, P.N);
    const uint32_t lbo_a = (uint32_t)npix * 16u, lbo_b = (uint32_t)P.N * 16u;
    for (int st = 0; st < P.n_stages; ++st) {
      const int slot = st & 1;
      tc::mbar_wait(&full_bar[slot], (uint32_t)((st >> 1) & 1));
      tc::tc_fence_after_sync();
      const uint32_t sa = tc::smem_u32(stage0 + (size_t)slot * stage_bytes);
      const uint32_t sb = sa + a_bytes;
      for (int m = 0; m < P.n_mt; ++m) {
        for (int tp = 0; tp < P.n_taps; ++tp) {
          const uint64_t ad = tc::smem_desc(sa + (uint32_t)(m * 128 + P.tap_off[tp]) * 16u, lbo_a, 128u);
          const uint64_t bd = tc::smem_desc(sb + (uint32_t)tp * 2u * lbo_b, lbo_b, 128u);
          tc::mma_tf32(tmem + (uint32_t)(m * P.N), ad, bd, idesc, (st > 0 || tp > 0) ? 1u : 0u);
        }
      }
      tc::mma_commit(&empty_bar[slot]);   // frees the smem slot once these MMAs have read it
    }
    tc::mma_commit(&done_bar);
  }
  __syncwarp();

  // ---------------- epilogue: 128 threads = 128 accumulator rows ----------------
  tc::mbar_wait(&done_bar, 0);
  tc::tc_fence_after_sync();
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const int plane = P.IY * P.IX;
  for (int m = 0; m < P.n_mt; ++m) {
    const int p = m * 128 + t;
    const int z = p / plane, rem = p - z * plane, y = rem / P.IX, x = rem - y * P.IX;
    const int gz = z0 + z, gy = y0 + y, gx = x0 + x;
    const bool valid = (z < P.TZ) && (y < P.TY) && (x < P.TX) && (gz < P.Dn) && (gy < P.Hn) && (gx < P.Wn);
    const size_t pix = ((size_t)gz * P.Hn + gy) * P.Wn + gx;
    for (int c0 = 0; c0 < P.N; c0 += 16) {
      float v[16];
      tc::tmem_ld16(trow + (uint32_t)(m * P.N + c0), v);
      tc::tmem_ld_wait();
      if (!valid) continue;
      if (P.mode == TC_PLAIN) {
        if (c0 >= P.cout) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j < P.cout) {
            v[j] += P.bias ? __ldg(P.bias + c0 + j) : 0.f;
            if (P.relu) v[j] = fmaxf(v[j], 0.f);
          }
        }
        float* o = P.out + pix * P.out_cstride + P.out_coff + c0;
        reinterpret_cast<float4*>(o)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(o)[1] = make_float4(v[4], v[5], v[6], v[7]);
        if (c0 + 8 < P.cout) {
          reinterpret_cast<float4*>(o)[2] = make_float4(v[8], v[9], v[10], v[11]);
          reinterpret_cast<float4*>(o)[3] = make_float4(v[12], v[13], v[14], v[15]);
        }
      } else if (P.mode == TC_HEAD) {   // feat_conv (8) + depth_conv (1), no bias
        float* o = P.out + pix * 8;
        reinterpret_cast<float4*>(o)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(o)[1] = make_float4(v[4], v[5], v[6], v[7]);
        P.out2[pix] = v[8];
      } else if (P.mode == TC_SINGLE) {
        P.out[pix] = v[0];
      } else {  // TC_DECONV: columns = parity * cout + co ; out[2b+e] = skip + (acc + bias)
        const int Ho = 2 * P.Hn, Wo = 2 * P.Wn;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = c0 + 8 * h;
          const int e = c / P.cout, co = c - e * P.cout;
          const size_t opix = ((size_t)(2 * gz + (e >> 2)) * Ho + (2 * gy + ((e >> 1) & 1))) * Wo + (2 * gx + (e & 1));
          const float* sk = P.skip + opix * P.cout + co;
          const float4 s0 = ldg4(sk), s1 = ldg4(sk + 4);
          const float4 b0 = ldg4(P.bias + co), b1 = ldg4(P.bias + co + 4);
          float* o = P.out + opix * P.cout + co;
          reinterpret_cast<float4*>(o)[0] = make_float4(s0.x + (v[8 * h + 0] + b0.x), s0.y + (v[8 * h + 1] + b0.y),
                                                        s0.z + (v[8 * h + 2] + b0.z), s0.w + (v[8 * h + 3] + b0.w));
          reinterpret_cast<float4*>(o)[1] = make_float4(s1.x + (v[8 * h + 4] + b1.x), s1.y + (v[8 * h + 5] + b1.y),
                                                        s1.z + (v[8 * h + 6] + b1.z), s1.w + (v[8 * h + 7] + b1.w));
        }
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, P.tmem_cols);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int tc_conv_launch(const TcConvLayer& L, const float* in, int Dn, int Hn, int Wn, const float* wpack, const float* bias,
                   const float* skip, float* out, float* out2, int out_cstride, int out_coff, cudaStream_t stream) {
  ENERF_REQUIRE(L.cin % 8 == 0 && L.cin >= 8, ENERF_EUNSUPPORTED, "tc_conv: cin %d must be a multiple of 8", L.cin);
  TcConvParams P;
  P.Dn = Dn, P.Hn = Hn, P.Wn = Wn;
  P.cout = L.cout, P.relu = L.relu, P.mode = L.mode;
  P.out_cstride = out_cstride, P.out_coff = out_coff;
  P.wpack = wpack, P.bias = bias, P.skip = skip, P.out = out, P.out2 = out2;
  P.n_stages = L.cin / 8;
  int n_real = (L.kind == 1) ? 8 * L.cout : L.cout;
  P.N = (n_real + 15) / 16 * 16;
  ENERF_REQUIRE(P.N <= 256, ENERF_EUNSUPPORTED, "tc_conv: N=%d > 256", P.N);
  if (L.mode == TC_DECONV) ENERF_REQUIRE(L.cout % 8 == 0 && skip && bias, ENERF_EINVAL, "tc_conv: deconv needs cout%%8==0, skip and bias");

  const int hz = (L.kind == 1) ? 1 : L.KD - 1, hy = (L.kind == 1) ? 1 : L.KH - 1, hx = hy;  // halo extents
  P.oz = (L.kind == 1) ? 0 : L.KD / 2, P.oy = (L.kind == 1) ? 0 : L.KH / 2, P.ox = P.oy;
  P.TX = 32;
  P.TY = (L.kind == 0 && L.KD == 1) ? 16 : 8;
  P.TZ = (L.kind == 0 && L.KD == 1) ? 1 : 2;
  if (P.TZ > Dn) P.TZ = Dn;
  if (P.TY > Hn) P.TY = Hn;
  size_t smem = 0;
  uint32_t stage_bytes = 0;
  for (;;) {
    P.IZ = P.TZ + hz, P.IY = P.TY + hy, P.IX = P.TX + hx;
    const int pmax = ((P.TZ - 1) * P.IY + (P.TY - 1)) * P.IX + P.TX - 1;
    P.n_mt = pmax / 128 + 1;
    // taps
    if (L.kind == 1) {
      P.n_taps = 8;
      for (int d = 0; d < 8; ++d) P.tap_off[d] = (((d >> 2) & 1) * P.IY + ((d >> 1) & 1)) * P.IX + (d & 1);
    } else {
      P.n_taps = L.KD * L.KH * L.KH;
      int i = 0;
      for (int kz = 0; kz < L.KD; ++kz)
        for (int ky = 0; ky < L.KH; ++ky)
          for (int kx = 0; kx < L.KH; ++kx) P.tap_off[i++] = (kz * P.IY + ky) * P.IX + kx;
    }
    const int npix = P.IZ * P.IY * P.IX;
    stage_bytes = ((uint32_t)npix * 32u + (uint32_t)P.n_taps * (uint32_t)P.N * 32u + 127u) & ~127u;
    const int max_off = P.tap_off[P.n_taps - 1];
    smem = 2 * (size_t)stage_bytes + (size_t)(P.n_mt * 128 + max_off + 8) * 16;
    if (P.n_mt * P.N <= 512 && smem <= 200 * 1024) break;
    if (P.TY > 2) P.TY /= 2;
    else if (P.TZ > 1) P.TZ /= 2;
    else ENERF_REQUIRE(false, ENERF_EUNSUPPORTED, "tc_conv: no tile fits (N=%d taps=%d)", P.N, P.n_taps);
  }
  uint32_t cols = 32;
  while ((int)cols < P.n_mt * P.N) cols <<= 1;
  P.tmem_cols = cols;

  PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode();
  ENERF_REQUIRE(encode != nullptr, ENERF_ECUDA, "tc_conv: cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap tmap;
  const cuuint64_t dims[5] = {4, (cuuint64_t)Wn, (cuuint64_t)Hn, (cuuint64_t)Dn, (cuuint64_t)(L.cin / 4)};
  const cuuint64_t strides[4] = {(cuuint64_t)L.cin * 4, (cuuint64_t)Wn * L.cin * 4, (cuuint64_t)Hn * Wn * L.cin * 4, 16};
  const cuuint32_t box[5] = {4, (cuuint32_t)P.IX, (cuuint32_t)P.IY, (cuuint32_t)P.IZ, 2};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(in), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ENERF_REQUIRE(cr == CUDA_SUCCESS, ENERF_ECUDA, "tc_conv: cuTensorMapEncodeTiled failed (%d) dims %dx%dx%dx%d box %dx%dx%d", (int)cr,
                Dn, Hn, Wn, L.cin, P.IZ, P.IY, P.IX);

  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tc_conv: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    smem_set = smem;
  }
  dim3 grid(ceil_div(Wn, P.TX), ceil_div(Hn, P.TY), ceil_div(Dn, P.TZ));
  tc_conv_kernel<<<grid, 128, smem, stream>>>(tmap, P);
  ENERF_CHECK_LAUNCH("tc_conv");
  return ENERF_OK;
}

}  // namespace enerf

extern "C" int enerf_tc_conv(int kind, int KD, int KH, int cin, int cout, int mode, int relu, const float* in, int D, int H, int W,
                             const float* wpack, const float* bias, const float* skip, float* out, float* out2, int out_cstride,
                             int out_coff, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(in && wpack && out, ENERF_EINVAL, "tc_conv: null pointer");
  ENERF_REQUIRE(kind == 0 || kind == 1, ENERF_EINVAL, "tc_conv: kind %d", kind);
  ENERF_REQUIRE((KD == 1 || KD == 3) && (KH == 1 || KH == 3), ENERF_EUNSUPPORTED, "tc_conv: kernel %dx%dx%d", KD, KH, KH);
  TcConvLayer L{kind, KD, KH, cin, cout, mode, relu};
  return tc_conv_launch(L, in, D, H, W, wpack, bias, skip, out, out2, out_cstride, out_coff, (cudaStream_t)stream);
}
