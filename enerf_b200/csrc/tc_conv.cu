// tc_conv.cu -- host side of the tcgen05 implicit-GEMM convolution (kernel + design: tc_conv.cuh):
// tile selection, TMA tensor-map encoding (cuTensorMapEncodeTiled through the runtime's driver
// entry point -- no link-time dependency on libcuda), launch, and the C-ABI test/diagnostic entry.
#include <cudaTypedefs.h>

#include "tc_conv.cuh"

namespace enerf {

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
