// tc_conv.cuh -- implicit-GEMM convolution on the tcgen05 tensor cores (TF32 operands, fp32
// accumulation in TMEM) for the channels-last CNNs of ENeRF: every stride-1 convolution of
// FeatureNet / (Min)CostRegNet, and the ConvTranspose3d layers in 8-phase sub-pixel form.
//
// Formulation.  A CTA owns a TZ x TY x TX tile of positions.  The tile PLUS its halo is staged in
// shared memory 8 input channels at a time as two planes [c4][z][y][x][4 floats] (16-byte cp.async
// with zero fill = the volume's zero padding).  In that layout "pixel p, channels 4c..4c+3"
// sits at  c*PLANE + p*16 bytes, which is exactly the K-major / no-swizzle UMMA operand layout
// (tc.cuh) with rows = consecutive LINEAR halo positions.  So the A operand of filter tap
// (kz,ky,kx) for the 128 positions p0..p0+127 is the SAME buffer with its start address advanced
// by ((kz*IY+ky)*IX+kx)*16 bytes: the 27 taps are 27 descriptors, no im2col copy exists anywhere.
// Rows whose linear position falls in the halo columns compute garbage and are dropped in the
// epilogue (73-84 % of the issued rows are real outputs; the tensor pipe is far from the limit).
//   K loop : Cin/8 stages, 2-deep ring (full/empty mbarriers); the stage's weights arrive by one
//            TMA bulk copy (cp.async.bulk) next to the activations.
//   MMA    : one thread issues n_stages * n_mtiles * n_taps tcgen05.mma (M=128, N=Cout padded to
//            16, K=8); accumulators of all M-tiles of the CTA live in TMEM (n_mtiles*N columns).
//   Stride 2 : the halo is staged as 2^d phase tiles X_r[q] = in[2q + r]; a tap k - pad = 2d + r reads
//            phase r at offset d, i.e. once more just a different operand start address.
//   Epilogue: 128 threads = 128 TMEM lanes; tcgen05.ld, bias / ReLU / skip add, coalesced
//            channels-last stores; for the transposed convolutions the 8 output parities are the
//            N dimension (N = 8*Cout) and the epilogue does the pixel shuffle.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace enerf {

enum TcConvMode { TC_PLAIN = 0, TC_HEAD = 1, TC_DECONV = 2, TC_SINGLE = 3 };

// ---- host side ------------------------------------------------------------------------------------
struct TcConvLayer {
  int kind;          // 0: conv (KD x KH x KH, pad K/2, stride 1), 1: transposed conv k3 s2 p1 op1
  int KD, KH;        // conv kernel extents (KD = 1 for 2-D layers)
  int cin, cout;
  int mode, relu;
  int stride = 1;    // 1 | 2 (kind 0 only).  Dn,Hn,Wn passed to tc_conv_launch are ALWAYS the output grid.
};

// FeatureNet's lateral fused into the following smooth convolution (tc_conv2.cu, PROD > 0): the convolution's 32-channel
// input  bilinear_x2(up_in) + (lat_b + lat_w^T lat_in)  is computed tile by tile in shared memory and never written to HBM.
struct TcLateral {
  int lat_cin;           // 8 (lat0: conv0) | 16 (lat1: conv1)
  const float* lat_in;   // (S,H,W,lat_cin) channels-last
  const float* lat_w;    // [lat_cin][32]
  const float* lat_b;    // [32]
  const float* up_in;    // (S,H/2,W/2,32)
  // optional second output of the fused launch: the (S,H,W,12) records [8 features | rgb * 0.5 + 0.5 | 0] the ray kernels gather
  // from (what enerf_pack_img_feat builds from feat_l2 and the NCHW source images) -- written by the same epilogue
  const float* rgb_src = nullptr;   // (S,3,H,W)
  float* packed_out = nullptr;      // (S,H,W,12)
};

// tc_conv2.cu: the persistent TMA-fed kernel.  Returns ENERF_OK when it launched the layer, 1 when the layer is
// not eligible (stride 2, weights beyond shared memory, ...) -- the caller then uses tc_conv.cu's kernel -- or an error.
bool tc_fold_rule(const TcConvLayer& L);     // which layers fold their kx taps into N (tc_conv.cu; packing.tc_fold_kx mirrors it)
int tc_fold_rule_level();
int tc_conv2_try_launch(const TcConvLayer& L, const float* in, int Dn, int Hn, int Wn, const float* wpack, const float* bias, const float* skip,
                        float* out, float* out2, int out_cstride, int out_coff, bool fold, cudaStream_t stream, const TcLateral* lat = nullptr);
int tc_conv2_impl();   // 0 auto, 1 tc_conv.cu only, 2 auto + stride-2 layers
bool tc_conv2_fuse_lateral();

int tc_conv_launch(const TcConvLayer& L, const float* in, int Dn, int Hn, int Wn, const float* wpack, const float* bias,
                   const float* skip, float* out, float* out2, int out_cstride, int out_coff, cudaStream_t stream);

}  // namespace enerf
