// tc_conv.cuh -- implicit-GEMM convolution on the tcgen05 tensor cores (TF32 operands, fp32
// accumulation in TMEM) for the channels-last CNNs of ENeRF: every stride-1 convolution of
// FeatureNet / (Min)CostRegNet, and the ConvTranspose3d layers in 8-phase sub-pixel form.
//
// Formulation.  A CTA owns a TZ x TY x TX tile of positions.  TMA (cp.async.bulk.tensor.5d) stages
// the tile PLUS its halo, 8 input channels at a time, as two planes [c4][z][y][x][4 floats]; the
// volume's zero padding is TMA's out-of-bounds fill.  In that layout "pixel p, channels 4c..4c+3"
// sits at  c*PLANE + p*16 bytes, which is exactly the K-major / no-swizzle UMMA operand layout
// (tc.cuh) with rows = consecutive LINEAR halo positions.  So the A operand of filter tap
// (kz,ky,kx) for the 128 positions p0..p0+127 is the SAME buffer with its start address advanced
// by ((kz*IY+ky)*IX+kx)*16 bytes: the 27 taps are 27 descriptors, no im2col copy exists anywhere.
// Rows whose linear position falls in the halo columns compute garbage and are dropped in the
// epilogue (73-84 % of the issued rows are real outputs; the tensor pipe is far from the limit).
//   K loop : Cin/8 stages, 2-deep TMA ring (full/empty mbarriers), weights of the stage arrive by
//            a 1-D bulk copy next to the activations.
//   MMA    : one thread issues n_stages * n_mtiles * n_taps tcgen05.mma (M=128, N=Cout padded to
//            16, K=8); accumulators of all M-tiles of the CTA live in TMEM (n_mtiles*N columns).
//   Epilogue: 128 threads = 128 TMEM lanes; tcgen05.ld, bias / ReLU / skip add, coalesced
//            channels-last stores; for the transposed convolutions the 8 output parities are the
//            N dimension (N = 8*Cout) and the epilogue does the pixel shuffle.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace enerf {

enum TcConvMode { TC_PLAIN = 0, TC_HEAD = 1, TC_DECONV = 2, TC_SINGLE = 3 };

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

// ---- host side ------------------------------------------------------------------------------------
struct TcConvLayer {
  int kind;          // 0: conv (KD x KH x KH, pad K/2, stride 1), 1: transposed conv k3 s2 p1 op1
  int KD, KH;        // conv kernel extents (KD = 1 for 2-D layers)
  int cin, cout;
  int mode, relu;
};

int tc_conv_launch(const TcConvLayer& L, const float* in, int Dn, int Hn, int Wn, const float* wpack, const float* bias,
                   const float* skip, float* out, float* out2, int out_cstride, int out_coff, cudaStream_t stream);

}  // namespace enerf
