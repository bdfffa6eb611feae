// tc_conv2.cu -- the persistent, warp-specialised, TMA-fed form of the tcgen05 implicit-GEMM convolution
// (stride-1 convolutions and the sub-pixel transposed convolutions of FeatureNet / (Min)CostRegNet whose
// weights fit in shared memory; everything else stays on tc_conv.cu's kernel).
//
// What changed against tc_conv.cu (same GEMM formulation: rows = linear halo positions, taps = operand
// start-address offsets, TF32 operands, fp32 accumulators in TMEM):
//   * The halo tile of a K-block (8 / 16 / 32 input channels of EVERY pixel) arrives by ONE TMA tensor-map box
//     {C, IX, IY, IZ} of the channels-last tensor, written as rows of C*4 bytes with SWIZZLE_{32,64,128}B -- the
//     K-major swizzled operand layout (csrc/tma.cuh).  Pixel p is operand row p, so a filter tap is still a
//     start-address offset (of whole rows), a K-step of 8 channels is +32 bytes inside the row, and the
//     volume's zero padding is the box's out-of-bounds fill.  (tests: tcgen05_swizzled_tma_operand_with_row_offsets)
//   * Persistent CTAs (static round-robin over tiles) with three roles:
//       warp 0            producer : waits for a free ring slot, issues the box of the next (tile, K-block)
//       warps 1..NMMA     MMA      : wait for the slot, issue taps x K-steps x M-tiles tcgen05.mma, commit -> slot free,
//                                    after the last K-block commit -> accumulator full
//       last 4 warps      epilogue : tcgen05.ld, bias / ReLU / fold shift / pixel shuffle + skip, coalesced stores,
//                                    then hand the accumulator back
//     The accumulators are double buffered in TMEM (2 x n_mt x N columns), so the epilogue of tile i runs under the
//     MMAs of tile i+1, and the ring keeps 2-4 boxes in flight.  The layer's weights are loaded once per CTA.
//   * kx folding (FOLD = 1: the three kx taps ride in N, 3x fewer MMAs) with a batched epilogue: the accumulator rows of 2-4 M-tiles
//     are read with one tcgen05.wait::ld, the two rows a warp needs from the next 32-row group travel through a small shared
//     exchange buffer (vector loads + selects, one named barrier per batch); one MMA warp, so the batch stays in registers.
//   * PROD > 0 (FeatureNet lat0 fused into smooth0): PROD computing producer warps build the 32-channel operand tile
//     (1x1 lateral + bilinear x2 + add) from TMA-staged source tiles; the epilogue can also write the ray kernels'
//     (feature | rgb) records.  All shared-memory traffic of the producer and of the exchange uses 32-bit shared-window addresses
//     (a pointer derived from the aligned dynamic-shared base is a GENERIC pointer to the compiler).
//   What the in-kernel timelines and ncu's source page said about each step: profiles/r2_conv2_timeline.md, r2_spin_wait.md.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tc_conv.cuh"
#include "tma.cuh"

namespace enerf {

struct TcConv2Params {
  int Dn, Hn, Wn;        // row grid (output grid of a conv; INPUT grid of a transposed conv)
  int TZ, TY, TX;        // tile
  int IZ, IY, IX;        // tile + halo
  int oz, oy, ox;        // halo origin = tile origin - (oz,oy,ox)
  int nx, ny, nz, n_tiles;
  int sz, sy, sx;        // stride per dimension (1|2): a stride-2 halo is staged as sz*sy*sx PHASE tiles X_r[q] = in[2q + r], each one
  int n_phases;          //   TMA box with element stride 2 starting at 2*(tile origin - o) + r; a tap k - pad = 2d + r reads phase r at offset d
  uint32_t phase_bytes;  // distance between phase tiles inside a slot (1024-aligned)
  int tap_off[27];       // operand start offsets of the taps in 16-byte units (phase * phase_bytes + rows * row_bytes) / 16
  int kbc, n_kb;         // channels per K-block (box channel extent: 8 | 16 | 32), K-blocks per tile (cin / kbc)
  int n_slots;           // ring depth
  uint32_t slot_bytes, box_bytes, w_bytes, xch_bytes;
  int N, n_mt, n_acc;    // MMA N, 128-row M-tiles per tile, accumulator stages (1 | 2)
  int d128z, d128y, d128x;   // halo-position step of one M-tile: 128 rows = d128z planes + d128y rows + d128x pixels
  int cout, relu, mode;
  int out_cstride, out_coff;
  uint32_t tmem_cols;
  const float* wpack;    // [cin/8][tap][2][N][4] TF32 (packing.pack_tc_conv / pack_tc_deconv)
  const float* bias;
  const float* skip;
  float* out;
  float* out2;
  // PROD > 0 (FeatureNet smooth0 with the lateral fused in): the 32-channel input tile is COMPUTED by PROD producer
  // warps instead of loaded: in[n,y,x,:] = bilinear_x2(up_in)[n,y,x,:] + (lat_b + lat_w^T lat_in[n,y,x,:]), zero outside the image
  const float* lat_in;   // (S,H,W,8) channels-last   (read through map_c0)
  const float* lat_w;    // [8][32]
  const float* lat_b;    // [32]
  const float* up_in;    // (S,H/2,W/2,32)            (read through map_f1)
  const float* rgb_src;  // (S,3,H,W) source images: with packed_out, the epilogue also writes [8 features | rgb * 0.5 + 0.5 | 0] records
  float* packed_out;     // (S,H,W,12)
  // the producer's SOURCE tiles arrive by TMA into a 2-deep ring behind the fold exchange: [c0 tile (IY x IX px x 32 B, SWIZZLE_32B)]
  // [f1 tile (FH x FW px x 128 B, SWIZZLE_128B)]
  int FW, FH;
  uint32_t src_slot_bytes, src_f1_off, src_bytes;   // slot stride, offset of the f1 tile inside a slot, bytes per slot's two boxes
  unsigned long long* dbg;                           // optional %globaltimer stamps of CTA 0 (enerf_tc_conv2_debug): [role 0..2][tile 0..15][8]
};

__device__ __forceinline__ unsigned long long gtime2() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// role 0 = producer (lane 0), 1 = MMA warp 0 (lane 0), 2 = epilogue (row 0)
#define C2_STAMP(role_, tile_, i_)                                                                  \
  do {                                                                                              \
    if (dbg && (tile_) < 16) dbg[((role_) * 16 + (tile_)) * 8 + (i_)] = gtime2();                   \
  } while (0)

__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
  return p != 0;
}

template <int NTAPS, int MODE, int FOLD, int NMMA, int PROD = 0>
__global__ void __launch_bounds__(32 * (4 + NMMA + 1 + PROD), PROD ? 1 : 2)      // two persistent CTAs per SM (one with the fused lateral)
    tc_conv2_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap map_c0, const __grid_constant__ CUtensorMap map_f1,
                    const TcConv2Params P) {
  constexpr int NP = 1 + PROD;          // producer warps: the TMA warp (+ PROD computing warps when the lateral is fused)
  constexpr int NPC = 32 * PROD;        // computing producer threads
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], accf_bar[2], acce_bar[2], w_bar, srcf_bar[2], srce_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // dynamic shared memory: [ring][weights][fold exchange x2]; the ring base is 1024-aligned (swizzle pattern phase)
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* w_s = ring + (size_t)P.n_slots * P.slot_bytes;
  float* xch0 = reinterpret_cast<float*>(w_s + ((P.w_bytes + 127u) & ~127u));

  if (t == 0) {
    for (int i = 0; i < 8; ++i) {
      tc::mbar_init(&full_bar[i], PROD != 0 ? NPC : 1);   // the producer's expect_tx arrival (+ the box's bytes) | every computing producer thread
      tc::mbar_init(&empty_bar[i], NMMA);    // one tcgen05.commit per MMA warp
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&accf_bar[i], NMMA);
      tc::mbar_init(&acce_bar[i], 128);      // every epilogue thread
    }
    tc::mbar_init(&w_bar, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&srcf_bar[i], 1);        // source boxes landed (expect_tx)
      tc::mbar_init(&srce_bar[i], NPC);      // every computing producer thread has read them
    }
    tc::fence_mbar_init();
  }
  if (warp == NP) tc::tmem_alloc(&tmem_base_s, P.tmem_cols);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t row_bytes = (uint32_t)P.kbc * 4u;

  if (PROD != 0 && warp == 0) {
    // ============================== source loader (fused lateral): conv0 tile + feat1_pre region of every tile, by TMA ==============================
    unsigned char* src0 = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(xch0) + P.xch_bytes + 1023) & ~(uintptr_t)1023);
    if (lane == 0) {
      tma::prefetch_desc(&map_c0);
      tma::prefetch_desc(&map_f1);
      tc::mbar_expect_tx(&w_bar, P.w_bytes);
      tc::tma_load_1d(w_s, P.wpack, P.w_bytes, &w_bar);
    }
    const int H = P.Hn, W = P.Wn, hi = H / 2, wi = W / 2;
    const float rh = (H > 1) ? (float)(hi - 1) / (float)(H - 1) : 0.f;
    const float rw = (W > 1) ? (float)(wi - 1) / (float)(W - 1) : 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
      const int bx = tile % P.nx, by = (tile / P.nx) % P.ny, n = tile / (P.nx * P.ny);
      const int x0 = bx * P.TX - P.ox, y0 = by * P.TY - P.oy;
      const int fy0 = (int)(rh * (float)max(y0, 0)), fx0 = (int)(rw * (float)max(x0, 0));     // first source row / column any pixel of the tile taps
      const int ss = it & 1;
      tc::mbar_wait(&srce_bar[ss], (uint32_t)(((it >> 1) & 1) ^ 1));
      if (lane == 0) {
        tc::mbar_expect_tx(&srcf_bar[ss], P.src_bytes);
        const uint32_t dst = tc::smem_u32(src0 + (size_t)ss * P.src_slot_bytes);
        tma::load_4d(dst, &map_c0, 0, x0, y0, n, &srcf_bar[ss]);
        tma::load_4d(dst + P.src_f1_off, &map_f1, 0, fx0, fy0, n, &srcf_bar[ss]);
      }
      __syncwarp();
    }
  } else if (PROD != 0 && warp < NP) {
    // ============================== computing producer (lateral 1x1 conv + bilinear x2 + add), sources in shared memory ==============================
    // thread = (halo pixel, g): channels [4g, 4g+4) and [16+4g, 16+4g+4) -- 16-byte chunks g and 4+g of the pixel's 128-byte
    // row, written at their SWIZZLE_128B position (chunk ^ (row & 7)); same arithmetic and order as lateral_upadd_kernel
    unsigned char* src0 = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(xch0) + P.xch_bytes + 1023) & ~(uintptr_t)1023);
    const int pt = t - 32;                              // 0..NPC-1
    const int g = pt & 3;
    float4 wreg[16];                                    // this thread's 8 x 8 slice of the lateral's weights
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      wreg[2 * k] = ldg4(P.lat_w + k * 32 + g * 4);
      wreg[2 * k + 1] = ldg4(P.lat_w + k * 32 + 16 + g * 4);
    }
    const float4 bias0 = ldg4(P.lat_b + g * 4), bias1 = ldg4(P.lat_b + 16 + g * 4);
    const int H = P.Hn, W = P.Wn, hi = H / 2, wi = W / 2;
    const float rh = (H > 1) ? (float)(hi - 1) / (float)(H - 1) : 0.f;
    const float rw = (W > 1) ? (float)(wi - 1) / (float)(W - 1) : 0.f;
    const int npix = P.IY * P.IX;
    // this thread's halo pixels: pp = (pt >> 2) + k * NPC / 4 (tile-invariant walk, no division per item)
    constexpr int dpp = NPC / 4;
    const int pp_first = pt >> 2, yy_first = pp_first / P.IX, xx_first = pp_first - yy_first * P.IX;
    const int dyy = dpp / P.IX, dxx = dpp - dyy * P.IX;
    const uint32_t ring_u = tc::smem_u32(ring), src0_u = tc::smem_u32(src0);
    int it = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++it) {
      const int bx = tile % P.nx, by = (tile / P.nx) % P.ny;
      const int x0 = bx * P.TX - P.ox, y0 = by * P.TY - P.oy;
      const int fy0 = (int)(rh * (float)max(y0, 0)), fx0 = (int)(rw * (float)max(x0, 0));
      const int slot = it % P.n_slots, ss = it & 1;
      unsigned long long* dbg = (P.dbg && blockIdx.x == 0 && pt == 0) ? P.dbg : nullptr;
      C2_STAMP(0, it, 0);
      tc::mbar_wait(&empty_bar[slot], (uint32_t)(((it / P.n_slots) & 1) ^ 1));       // operand slot free (its MMAs are done)
      C2_STAMP(0, it, 1);
      tc::mbar_wait(&srcf_bar[ss], (uint32_t)((it >> 1) & 1));                       // this tile's sources have landed
      C2_STAMP(0, it, 2);
      // 32-bit shared-window addresses + ld/st.shared: the byte pointers derived from the aligned dynamic-shared base had lost
      // their address space (generic LD/ST with 64-bit address arithmetic), and the per-item pixel -> (row, column) division is
      // replaced by an incremental walk: ncu counted ~330 issued instructions per item for ~100 FMAs (profiles/r2_spin_wait.md).
      const uint32_t dst_u = ring_u + (uint32_t)slot * P.slot_bytes;
      const uint32_t c0_u = src0_u + (uint32_t)ss * P.src_slot_bytes, f1_u = c0_u + P.src_f1_off;
      int pp = pp_first, yy = yy_first, xx = xx_first;
      for (; pp < npix; pp += dpp) {
        const int y = y0 + yy, x = x0 + xx;
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
          float acc[8] = {bias0.x, bias0.y, bias0.z, bias0.w, bias1.x, bias1.y, bias1.z, bias1.w};
          const uint32_t crow = c0_u + (uint32_t)pp * 32u;             // SWIZZLE_32B: 16-byte chunk j of row pp sits at j ^ ((pp >> 2) & 1)
          const uint32_t csw = (uint32_t)((pp >> 2) & 1);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 v = lds128(crow + (((uint32_t)q ^ csw) << 4));
            const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int jx = 0; jx < 4; ++jx) {
              const float4 w0 = wreg[2 * (4 * q + jx)], w1 = wreg[2 * (4 * q + jx) + 1];
              acc[0] = fmaf(xv[jx], w0.x, acc[0]), acc[1] = fmaf(xv[jx], w0.y, acc[1]), acc[2] = fmaf(xv[jx], w0.z, acc[2]), acc[3] = fmaf(xv[jx], w0.w, acc[3]);
              acc[4] = fmaf(xv[jx], w1.x, acc[4]), acc[5] = fmaf(xv[jx], w1.y, acc[5]), acc[6] = fmaf(xv[jx], w1.z, acc[6]), acc[7] = fmaf(xv[jx], w1.w, acc[7]);
            }
          }
          // bilinear x2 with align_corners=True (ATen upsample_bilinear2d lambdas); taps from the staged feat1_pre region
          const float h1r = rh * (float)y, w1r = rw * (float)x;
          const int h1 = (int)h1r, w1 = (int)w1r;
          const int h1p = (h1 < hi - 1) ? 1 : 0, w1p = (w1 < wi - 1) ? 1 : 0;
          const float h1l = h1r - (float)h1, h0l = 1.f - h1l, w1l = w1r - (float)w1, w0l = 1.f - w1l;
          const int r00 = (h1 - fy0) * P.FW + (w1 - fx0), r01 = r00 + w1p, r10 = r00 + h1p * P.FW, r11 = r10 + w1p;
          auto tap = [&](int rr, int chunk) {       // SWIZZLE_128B: chunk c of row rr sits at c ^ (rr & 7)
            return lds128(f1_u + (uint32_t)rr * 128u + ((uint32_t)(chunk ^ (rr & 7)) << 4));
          };
          float up[8];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 a = tap(r00, g + 4 * q), b = tap(r01, g + 4 * q), c = tap(r10, g + 4 * q), dd = tap(r11, g + 4 * q);
            up[4 * q + 0] = h0l * (w0l * a.x + w1l * b.x) + h1l * (w0l * c.x + w1l * dd.x);
            up[4 * q + 1] = h0l * (w0l * a.y + w1l * b.y) + h1l * (w0l * c.y + w1l * dd.y);
            up[4 * q + 2] = h0l * (w0l * a.z + w1l * b.z) + h1l * (w0l * c.z + w1l * dd.z);
            up[4 * q + 3] = h0l * (w0l * a.w + w1l * b.w) + h1l * (w0l * c.w + w1l * dd.w);
          }
          // reference order: interpolate(x) + lateral(y)   (feature_net.py:25)
          o0 = make_float4(up[0] + acc[0], up[1] + acc[1], up[2] + acc[2], up[3] + acc[3]);
          o1 = make_float4(up[4] + acc[4], up[5] + acc[5], up[6] + acc[6], up[7] + acc[7]);
        }
        const uint32_t rowp = dst_u + (uint32_t)pp * 128u;
        sts128(rowp + ((uint32_t)(g ^ (pp & 7)) << 4), o0);
        sts128(rowp + ((uint32_t)((4 + g) ^ (pp & 7)) << 4), o1);
        xx += dxx, yy += dyy;
        if (xx >= P.IX) xx -= P.IX, ++yy;
      }
      mbar_arrive1(&srce_bar[ss]);             // the source slot may be refilled
      tc::fence_proxy_async();                 // my part of the tile -> visible to the tensor core
      mbar_arrive1(&full_bar[slot]);
      C2_STAMP(0, it, 3);                      // operand tile written
    }
  } else if (warp == 0) {
    // ============================== producer ==============================
    if (lane == 0) {
      tma::prefetch_desc(&map);
      tc::mbar_expect_tx(&w_bar, P.w_bytes);
      tc::tma_load_1d(w_s, P.wpack, P.w_bytes, &w_bar);
    }
    int it = 0, tcount = 0;
    unsigned long long* dbg = (P.dbg && blockIdx.x == 0 && lane == 0) ? P.dbg : nullptr;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++tcount) {
      const int bx = tile % P.nx, by = (tile / P.nx) % P.ny, bz = tile / (P.nx * P.ny);
      const int x0 = bx * P.TX - P.ox, y0 = by * P.TY - P.oy, z0 = bz * P.TZ - P.oz;
      C2_STAMP(0, tcount, 0);
      for (int kb = 0; kb < P.n_kb; ++kb, ++it) {
        const int slot = it % P.n_slots;
        tc::mbar_wait(&empty_bar[slot], (uint32_t)(((it / P.n_slots) & 1) ^ 1));
        if (kb == 0) C2_STAMP(0, tcount, 1);       // first slot free
        if (lane == 0) {
          tc::mbar_expect_tx(&full_bar[slot], P.box_bytes * (uint32_t)P.n_phases);
          const uint32_t dst = tc::smem_u32(ring + (size_t)slot * P.slot_bytes);
          for (int ph = 0; ph < P.n_phases; ++ph) {
            const int rx = ph % P.sx, ry = (ph / P.sx) % P.sy, rz = ph / (P.sx * P.sy);
            tma::load_4d(dst + (uint32_t)ph * P.phase_bytes, &map, kb * P.kbc, P.sx * x0 + rx, P.sy * y0 + ry, P.sz * z0 + rz, &full_bar[slot]);
          }
        }
        __syncwarp();
      }
      C2_STAMP(0, tcount, 2);                       // boxes issued
    }
  } else if (warp < NP + NMMA) {
    // ============================== MMA issue (converged warp, one elected lane issues) ==============================
    const int mw = warp - NP;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = tc::idesc_tf32(128, P.N);
    const uint64_t a_hi = tma::smem_desc_swz(0, row_bytes, 0), b_hi = tc::smem_desc(0, (uint32_t)P.N * 16u, 128u);
    const uint32_t w0 = tc::smem_u32(w_s) >> 4, b_step = ((uint32_t)P.N * 32u) >> 4;      // one (K-stage, tap) B block = N*32 bytes
    const uint32_t m_step = (128u * row_bytes) >> 4;
    const int ksteps = P.kbc / 8;
    tc::mbar_wait(&w_bar, 0);
    int it = 0, ti = 0;
    unsigned long long* dbg = (P.dbg && blockIdx.x == 0 && mw == 0 && lane == 0) ? P.dbg : nullptr;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++ti) {
      const int acc = ti % P.n_acc;
      C2_STAMP(1, ti, 0);
      tc::mbar_wait(&acce_bar[acc], (uint32_t)(((ti / P.n_acc) & 1) ^ 1));      // the epilogue has drained this accumulator
      tc::tc_fence_after_sync();
      C2_STAMP(1, ti, 1);                           // accumulator free
      const uint32_t td0 = tmem_u + (uint32_t)(acc * P.n_mt * P.N);
      for (int kb = 0; kb < P.n_kb; ++kb, ++it) {
        const int slot = it % P.n_slots;
        tc::mbar_wait(&full_bar[slot], (uint32_t)((it / P.n_slots) & 1));
        tc::tc_fence_after_sync();
        if (kb == 0) C2_STAMP(1, ti, 2);            // first K-block landed
        const uint32_t a0 = tc::smem_u32(ring + (size_t)slot * P.slot_bytes) >> 4;
        for (int m = mw; m < P.n_mt; m += NMMA) {
          const uint32_t td = td0 + (uint32_t)(m * P.N);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t am = a0 + (uint32_t)m * m_step + (uint32_t)ks * 2u;
            const uint32_t bk = w0 + (uint32_t)((kb * ksteps + ks) * NTAPS) * b_step;
#pragma unroll
            for (int tp = 0; tp < NTAPS; ++tp)
              tc::mma_tf32_elect(td, a_hi | (uint64_t)((am + (uint32_t)P.tap_off[tp]) & 0x3FFFu), b_hi | (uint64_t)((bk + (uint32_t)tp * b_step) & 0x3FFFu), idesc,
                                 (kb > 0 || ks > 0 || tp > 0) ? 1u : 0u);
          }
        }
        tc::mma_commit_elect(&empty_bar[slot]);                 // the slot is free once these MMAs have read it
        if (kb + 1 == P.n_kb) tc::mma_commit_elect(&accf_bar[acc]);
        __syncwarp();
      }
      C2_STAMP(1, ti, 3);                           // all MMAs of the tile issued
    }
  } else {
    // ============================== epilogue: 128 threads = 128 accumulator rows ==============================
    // Measured with the in-kernel stamps (profiles/r2_conv2_timeline.md): with ONE epilogue warp per scheduler the epilogue is a
    // latency chain, not a throughput problem -- so everything tile-invariant is hoisted (the row -> (z,y,x) decomposition is
    // advanced incrementally, no integer divisions per M-tile; the bias lives in registers), and the TMEM reads of several
    // M-tiles share one tcgen05.wait::ld.
    const int g = warp & 3;                                   // TMEM lane group this warp may read
    const int r = g * 32 + lane;                              // row of the M-tile
    const uint32_t trow0 = tmem + ((uint32_t)(g * 32) << 16);
    const int plane = P.IY * P.IX;
    const uint32_t xch0_u = tc::smem_u32(xch0);
    struct RowPos { int z, y, x; };
    const RowPos first = {r / plane, (r % plane) / P.IX, r % P.IX};                 // halo position of accumulator row r of M-tile 0
    auto advance = [&](RowPos& p) {                                                 // ... and of the same row one M-tile on
      p.x += P.d128x, p.y += P.d128y, p.z += P.d128z;
      if (p.x >= P.IX) p.x -= P.IX, ++p.y;
      if (p.y >= P.IY) p.y -= P.IY, ++p.z;
    };
    float bias_r[8];                                          // folded PLAIN layers (8 output channels): the bias lives in registers
#pragma unroll
    for (int c = 0; c < 8; ++c) bias_r[c] = (FOLD && MODE == TC_PLAIN) ? __ldg(P.bias + c) : 0.f;
    int ti = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++ti) {
      const int acc = ti % P.n_acc;
      const int bx = tile % P.nx, by = (tile / P.nx) % P.ny, bz = tile / (P.nx * P.ny);
      const int x0 = bx * P.TX, y0 = by * P.TY, z0 = bz * P.TZ;
      unsigned long long* dbg = (P.dbg && blockIdx.x == 0 && r == 0) ? P.dbg : nullptr;
      C2_STAMP(2, ti, 0);
      tc::mbar_wait(&accf_bar[acc], (uint32_t)((ti / P.n_acc) & 1));
      tc::tc_fence_after_sync();
      C2_STAMP(2, ti, 1);                           // accumulator full (all MMAs of the tile complete)
      const uint32_t trow = trow0 + (uint32_t)(acc * P.n_mt * P.N);
      const uint32_t xch_u = xch0_u + (uint32_t)(ti & 1) * (P.xch_bytes / 2);      // [n_mt*4 + 1][2 rows][N] floats, double buffered (shared-window address)
      RowPos pos = first;                            // advanced once per M-tile, in order
      auto locate = [&](const RowPos& p, size_t& pix, int& gz, int& gy, int& gx) -> bool {
        gz = z0 + p.z, gy = y0 + p.y, gx = x0 + p.x;
        pix = ((size_t)gz * P.Hn + gy) * P.Wn + gx;
        return (p.z < P.TZ) && (p.y < P.TY) && (p.x < P.TX) && (gz < P.Dn) && (gy < P.Hn) && (gx < P.Wn);
      };
      auto release_acc = [&]() {       // after the last TMEM read of this accumulator: hand it back to the MMA warps
        tc::tc_fence_before_sync();
        mbar_arrive1(&acce_bar[acc]);
        C2_STAMP(2, ti, 2);            // accumulator released
      };
      auto emit8 = [&](const float* vv, float* dst, const float* b8, float* dst2 = nullptr) {
        float4 o0 = make_float4(vv[0] + b8[0], vv[1] + b8[1], vv[2] + b8[2], vv[3] + b8[3]);
        float4 o1 = make_float4(vv[4] + b8[4], vv[5] + b8[5], vv[6] + b8[6], vv[7] + b8[7]);
        if (P.relu) {
          o0 = make_float4(fmaxf(o0.x, 0.f), fmaxf(o0.y, 0.f), fmaxf(o0.z, 0.f), fmaxf(o0.w, 0.f));
          o1 = make_float4(fmaxf(o1.x, 0.f), fmaxf(o1.y, 0.f), fmaxf(o1.z, 0.f), fmaxf(o1.w, 0.f));
        }
        reinterpret_cast<float4*>(dst)[0] = o0;
        reinterpret_cast<float4*>(dst)[1] = o1;
        if (dst2 != nullptr) {       // the same 8 channels as the head of a wider record
          reinterpret_cast<float4*>(dst2)[0] = o0;
          reinterpret_cast<float4*>(dst2)[1] = o1;
        }
      };
      // fused lateral only: second output = the ray kernels' (feature | rgb) record of this pixel (same arithmetic as
      // pack_img_feat_kernel: rgb * 0.5 + 0.5 from the NCHW source image of view gz)
      auto packed_dst = [&](size_t pix, int gz, int gy, int gx) -> float* {
        if constexpr (PROD != 0) {
          if (P.packed_out == nullptr) return nullptr;
          float* d = P.packed_out + pix * 12;
          const size_t hw = (size_t)P.Hn * P.Wn;
          const float* sp = P.rgb_src + (size_t)gz * 3 * hw + (size_t)gy * P.Wn + gx;
          reinterpret_cast<float4*>(d)[2] = make_float4(__ldg(sp) * 0.5f + 0.5f, __ldg(sp + hw) * 0.5f + 0.5f, __ldg(sp + 2 * hw) * 0.5f + 0.5f, 0.f);
          return d;
        } else {
          (void)pix, (void)gz, (void)gy, (void)gx;
          return nullptr;
        }
      };
      auto store8 = [&](const float* v, float* dst, const float* bias8, float* dst2 = nullptr) {      // bias from global memory (L1 hit)
        const float4 b0v = ldg4(bias8), b1v = ldg4(bias8 + 4);
        const float b8[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
        emit8(v, dst, b8, dst2);
      };
      if constexpr (FOLD) {
        // columns kx*C + co (C = cout | 9 | 1): out[q] = P[q][0:C] + P[q+1][C:2C] + P[q+2][2C:3C].  Rows q+1, q+2 sit in the next
        // lanes (shuffles); the two last lanes of a 32-row group need rows 0, 1 of the next group, which every warp publishes
        // through `xch`.  A batch of NB M-tiles (+ the first of the next batch, only for its rows 0, 1) is read with ONE wait and
        // exchanged with ONE barrier; the accumulator goes back to the MMA warps as soon as the last batch sits in registers.
        constexpr int NC = (MODE == TC_PLAIN) ? 24 : (MODE == TC_HEAD) ? 32 : 8;   // columns read per row (3C: 24 / 27 / 3)
        constexpr int NB = (MODE == TC_PLAIN) ? 3 : (MODE == TC_HEAD) ? 2 : 4;    // M-tiles per batch
        const int XS = P.N;                                                        // row stride of the exchange buffer
        auto ldrow = [&](int m, float* v) {
          const uint32_t a = trow + (uint32_t)(m * P.N);
          if constexpr (NC == 24) {
            tc::tmem_ld16(a, v);
            tc::tmem_ld8(a + 16, v + 16);
          } else if constexpr (NC == 32) {
            tc::tmem_ld32(a, v);
          } else {
            tc::tmem_ld8(a, v);
          }
        };
        auto publish = [&](int m, const float* v) {      // lanes 0, 1: rows 0, 1 of this warp's 32-row group of M-tile m
          const uint32_t d = xch_u + (uint32_t)(((m * 4 + g) * 2 + lane) * XS) * 4u;
#pragma unroll
          for (int c4 = 0; c4 < NC / 4; ++c4) sts128(d + 16u * c4, make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]));
        };
        auto outputs = [&](int m, const float* v) {
          size_t pix;
          int gz, gy, gx;
          const bool valid = locate(pos, pix, gz, gy, gx);
          advance(pos);
          // rows 0, 1 of the next 32-row group.  EVERY lane loads (uniform row 0 / lane-selected row, 16-byte vectors) and the two
          // last lanes SELECT: the per-element `if (lane >= 30) x = smem[..]` this replaces compiled to 16 divergent
          // branch regions per M-tile and made the fold epilogue ~1.3 us per M-tile (profiles/r2_conv2_timeline.md).
          const uint32_t row0 = xch_u + (uint32_t)((m * 4 + g + 1) * 2 * XS) * 4u;
          const uint32_t rowc = row0 + ((lane == 31) ? (uint32_t)XS * 4u : 0u);            // row (lane - 30) for lanes 30, 31
          auto ld4 = [](uint32_t q, float* d) {
            const float4 t4 = lds128(q);
            d[0] = t4.x, d[1] = t4.y, d[2] = t4.z, d[3] = t4.w;
          };
          if constexpr (MODE == TC_PLAIN) {      // C = 8 (the rule folds 8-channel layers only): columns [0,8) [8,16) [16,24)
            float nb[8], nc[8], a[8];
            ld4(row0 + 32, nb), ld4(row0 + 48, nb + 4);
            ld4(rowc + 64, nc), ld4(rowc + 80, nc + 4);
#pragma unroll
            for (int jx = 0; jx < 8; ++jx) {
              float b1 = __shfl_down_sync(0xffffffffu, v[8 + jx], 1);
              float c2 = __shfl_down_sync(0xffffffffu, v[16 + jx], 2);
              b1 = (lane == 31) ? nb[jx] : b1;
              c2 = (lane >= 30) ? nc[jx] : c2;
              a[jx] = (v[jx] + b1) + c2;
            }
            if (valid) emit8(a, P.out + pix * P.out_cstride + P.out_coff, bias_r, packed_dst(pix, gz, gy, gx));
          } else if constexpr (MODE == TC_HEAD) {   // columns kx*9 + co: 8 feat + 1 prob per kx
            float nb[12], nc[12], rr[9];            // columns 8..19 of row 0 (9..17 used), 16..27 of the selected row (18..26 used)
            ld4(row0 + 32, nb), ld4(row0 + 48, nb + 4), ld4(row0 + 64, nb + 8);
            ld4(rowc + 64, nc), ld4(rowc + 80, nc + 4), ld4(rowc + 96, nc + 8);
#pragma unroll
            for (int jx = 0; jx < 9; ++jx) {
              float b1 = __shfl_down_sync(0xffffffffu, v[9 + jx], 1);
              float c2 = __shfl_down_sync(0xffffffffu, v[18 + jx], 2);
              b1 = (lane == 31) ? nb[1 + jx] : b1;
              c2 = (lane >= 30) ? nc[2 + jx] : c2;
              rr[jx] = (v[jx] + b1) + c2;
            }
            if (valid) {
              float4* o = reinterpret_cast<float4*>(P.out + pix * 8);
              o[0] = make_float4(rr[0], rr[1], rr[2], rr[3]);
              o[1] = make_float4(rr[4], rr[5], rr[6], rr[7]);
              P.out2[pix] = rr[8];
            }
          } else {                                  // TC_SINGLE: columns 0, 1, 2
            float nb[4], nc[4];
            ld4(row0, nb), ld4(rowc, nc);
            float b1 = __shfl_down_sync(0xffffffffu, v[1], 1);
            float c2 = __shfl_down_sync(0xffffffffu, v[2], 2);
            b1 = (lane == 31) ? nb[1] : b1;
            c2 = (lane >= 30) ? nc[2] : c2;
            if (valid) P.out[pix] = (v[0] + b1) + c2;
          }
        };
        for (int m0 = 0; m0 < P.n_mt; m0 += NB) {
          float v[NB][NC], la[NC];
          const int nb = min(NB, P.n_mt - m0);
          const bool look = m0 + NB < P.n_mt;        // another batch follows: fetch its first M-tile's rows 0, 1 now
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < nb) ldrow(m0 + b, v[b]);
          if (look) ldrow(m0 + NB, la);
          tc::tmem_ld_wait();
          if (!look) release_acc();
          if (lane < 2) {
#pragma unroll
            for (int b = 0; b < NB; ++b)
              if (b < nb && (b > 0 || m0 == 0)) publish(m0 + b, v[b]);    // (a batch's first M-tile was published as the look-ahead)
            if (look) publish(m0 + NB, la);
          }
          epi_bar_sync();
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b < nb) outputs(m0 + b, v[b]);
        }
      } else if (MODE == TC_PLAIN && (P.cout == 8 || P.cout == 16)) {
        // narrow layers: the TMEM read latency is amortised over several M-tiles per wait (4 x 8 or 2 x 16 columns)
        auto put = [&](const float* vv, int nch) {
          size_t pix;
          int gz, gy, gx;
          const bool valid = locate(pos, pix, gz, gy, gx);
          advance(pos);
          if (valid) {
            float* dst = P.out + pix * P.out_cstride + P.out_coff;
            store8(vv, dst, P.bias, (nch == 8) ? packed_dst(pix, gz, gy, gx) : nullptr);
            if (nch == 16) store8(vv + 8, dst + 8, P.bias + 8);
          }
        };
        if (P.cout == 8) {
          for (int m0 = 0; m0 < P.n_mt; m0 += 4) {
            float v0[8], v1[8], v2[8], v3[8];
            const int nb = min(4, P.n_mt - m0);
            tc::tmem_ld8(trow + (uint32_t)(m0 * P.N), v0);
            if (nb > 1) tc::tmem_ld8(trow + (uint32_t)((m0 + 1) * P.N), v1);
            if (nb > 2) tc::tmem_ld8(trow + (uint32_t)((m0 + 2) * P.N), v2);
            if (nb > 3) tc::tmem_ld8(trow + (uint32_t)((m0 + 3) * P.N), v3);
            tc::tmem_ld_wait();
            if (m0 + nb == P.n_mt) release_acc();
            put(v0, 8);
            if (nb > 1) put(v1, 8);
            if (nb > 2) put(v2, 8);
            if (nb > 3) put(v3, 8);
          }
        } else {
          for (int m0 = 0; m0 < P.n_mt; m0 += 2) {
            float v0[16], v1[16];
            const int nb = min(2, P.n_mt - m0);
            tc::tmem_ld16(trow + (uint32_t)(m0 * P.N), v0);
            if (nb > 1) tc::tmem_ld16(trow + (uint32_t)((m0 + 1) * P.N), v1);
            tc::tmem_ld_wait();
            if (m0 + nb == P.n_mt) release_acc();
            put(v0, 16);
            if (nb > 1) put(v1, 16);
          }
        }
      } else {
        for (int m = 0; m < P.n_mt; ++m) {
          size_t pix;
          int gz, gy, gx;
          const bool valid = locate(pos, pix, gz, gy, gx);
          advance(pos);
          const bool last = (m + 1 == P.n_mt);
          if constexpr (MODE == TC_PLAIN) {        // 32 | 64 output channels: 32 columns per load, one wait for all of them
            float* dst = P.out + pix * P.out_cstride + P.out_coff;
            const uint32_t tcol = trow + (uint32_t)(m * P.N);
            float v[64];
            tc::tmem_ld32(tcol, v);
            if (P.cout == 64) tc::tmem_ld32(tcol + 32, v + 32);
            tc::tmem_ld_wait();
            if (last) release_acc();
            if (valid) {
#pragma unroll
              for (int c0 = 0; c0 < 64; c0 += 8)
                if (c0 < P.cout) store8(v + c0, dst + c0, P.bias + c0);
            }
          } else if constexpr (MODE == TC_HEAD) {          // feat_conv (8) + depth_conv (1), no bias
            float v[16];
            tc::tmem_ld16(trow + (uint32_t)(m * P.N), v);
            tc::tmem_ld_wait();
            if (last) release_acc();
            if (valid) {
              float4* o = reinterpret_cast<float4*>(P.out + pix * 8);
              o[0] = make_float4(v[0], v[1], v[2], v[3]);
              o[1] = make_float4(v[4], v[5], v[6], v[7]);
              P.out2[pix] = v[8];
            }
          } else if constexpr (MODE == TC_SINGLE) {
            float v[8];
            tc::tmem_ld8(trow + (uint32_t)(m * P.N), v);
            tc::tmem_ld_wait();
            if (last) release_acc();
            if (valid) P.out[pix] = v[0];
          } else {  // TC_DECONV: columns = parity * cout + co ; out[2b+e] = skip + (acc + bias); 64 columns per wait
            const int Ho = 2 * P.Hn, Wo = 2 * P.Wn;
            for (int cb = 0; cb < P.N; cb += 64) {
              float v[64];
              tc::tmem_ld32(trow + (uint32_t)(m * P.N + cb), v);
              tc::tmem_ld32(trow + (uint32_t)(m * P.N + cb + 32), v + 32);       // N = 8 * cout is a multiple of 64
              tc::tmem_ld_wait();
              if (last && cb + 64 >= P.N) release_acc();
              if (valid) {
#pragma unroll
                for (int c8 = 0; c8 < 64; c8 += 8) {
                  const int c = cb + c8;
                  const int e = c / P.cout, co = c - e * P.cout;
                  const size_t opix = ((size_t)(2 * gz + (e >> 2)) * Ho + (2 * gy + ((e >> 1) & 1))) * Wo + (2 * gx + (e & 1));
                  const float* sk = P.skip + opix * P.cout + co;
                  const float4 s0 = ldg4(sk), s1 = ldg4(sk + 4);
                  const float4 b0v = ldg4(P.bias + co), b1v = ldg4(P.bias + co + 4);
                  float4* o = reinterpret_cast<float4*>(P.out + opix * P.cout + co);
                  o[0] = make_float4(s0.x + (v[c8 + 0] + b0v.x), s0.y + (v[c8 + 1] + b0v.y), s0.z + (v[c8 + 2] + b0v.z), s0.w + (v[c8 + 3] + b0v.w));
                  o[1] = make_float4(s1.x + (v[c8 + 4] + b1v.x), s1.y + (v[c8 + 5] + b1v.y), s1.z + (v[c8 + 6] + b1v.z), s1.w + (v[c8 + 7] + b1v.w));
                }
              }
            }
          }
        }
      }
      C2_STAMP(2, ti, 3);              // outputs of the tile stored
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == NP) tc::tmem_dealloc(tmem, P.tmem_cols);
}

// ---- host side ------------------------------------------------------------------------------------
// FeatureNet: lat0 (1x1 conv + bilinear x2 + add) computed inside smooth0's producer warps.  First version (taps gathered from
// global memory by the producer threads): the producer was the bottleneck, 304 us against 84 + 58 us for the two kernels.
// Second version (conv0 / feat1_pre source tiles staged by TMA, taps and the 1x1 from shared memory): feature_net 0.361 -> 0.335 ms,
// frame 1.130 -> 1.114 ms, bit-identical features (profiles/r2_frame_ab.md) -> on by default.
static int g_fuse_lateral = 1;
static int g_impl = 0;          // 0 auto (every eligible layer incl. stride 2), 1 force tc_conv.cu's kernel, 2 = 0, 3 auto without the stride-2 layers
static bool stride2_enabled() { return g_impl != 3; }
// One issuing warp sustains one M=128,K=8 MMA per ~91 cycles, the tensor pipe takes one per ~46 (profiles/r2_mma_bench2.md):
// two issuing warps x two persistent CTAs per SM keep it fed (per-layer A/B: profiles/r2_conv2_sweep.md).
static int g_nmma = 2;          // MMA-issuing warps per CTA (1 | 2)
static int g_ctas_per_sm = 2;   // persistent CTAs per SM (1 | 2)
static int g_prod_warps = 6;    // computing producer warps of the fused lateral (4 | 6 | 8): 6 = 12 warps per CTA, 166 registers, no spill
static unsigned long long* g_conv2_dbg = nullptr;
static int g_conv2_dbg_lat_only = 0;
static int g_tune2_tz = 0, g_tune2_ty = 0, g_tune2_fold = -1, g_tune2_kbc = 0, g_tune2_slots = 0;

int tc_conv2_impl() { return g_impl; }
bool tc_conv2_fuse_lateral() { return g_impl != 1 && g_fuse_lateral != 0; }

// returns ENERF_OK when launched, 1 when the layer is not eligible (caller falls back to tc_conv_launch)
// Host geometry of a launch, free of CUDA calls (so tests/test_host_cpu.py can emulate the kernel from it on a CPU-only
// box through enerf_tc_conv2_plan): tile, halo, taps, K-block, ring, accumulators.  Returns 0, or 1 = not eligible.
static int tc_conv2_plan(const TcConvLayer& L, int Dn, int Hn, int Wn, bool fold_default, bool lat, bool has_bias, bool has_skip, int n_sm,
                         TcConv2Params& P, int& n_taps_out, bool& fold_out, int lat_cin = 8) {
  const int stride = (L.kind == 0) ? L.stride : 1;
  if (lat && lat_cin != 8) return 1;     // instantiated: lat0 (8 channels) in front of the 3x3 (kx taps folded or not)
  if (lat && !(L.kind == 0 && L.KD == 1 && L.KH == 3 && L.cin == 32 && stride == 1 && L.mode == TC_PLAIN && Hn % 2 == 0 && Wn % 2 == 0))
    return 1;
  if (L.cin % 8 != 0 || L.cin > 64 || (stride != 1 && stride != 2)) return 1;
  if (stride == 2 && (fold_default || !stride2_enabled())) return 1;
  P.Dn = Dn, P.Hn = Hn, P.Wn = Wn;
  P.cout = L.cout, P.relu = L.relu, P.mode = L.mode;
  const bool skip = has_skip, bias = has_bias;
  bool fold = fold_default;      // the weights are packed for the default rule (packing.tc_fold_kx): keep it
  const int n_real = (L.kind == 1) ? 8 * L.cout : fold ? 3 * L.cout : L.cout;
  P.N = (n_real + 15) / 16 * 16;
  if (P.N > 256) return 1;
  const int n_taps = (L.kind == 1) ? 8 : fold ? L.KD * L.KH : L.KD * L.KH * L.KH;
  P.w_bytes = (uint32_t)(L.cin / 8) * (uint32_t)n_taps * (uint32_t)P.N * 32u;
  if (P.w_bytes > 72 * 1024) return 1;
  if (L.mode == TC_DECONV && !(L.cout % 8 == 0 && skip && bias)) return 1;
  if (L.mode == TC_PLAIN && !(L.cout % 8 == 0 && bias)) return 1;
  // epilogue shapes the kernel is written for: one TMEM load per M-tile of 8 / 16 / 32 / 64 channels; kx-folded PLAIN layers have 8
  if (L.mode == TC_PLAIN && L.kind == 0 && !(fold ? L.cout == 8 : (L.cout == 8 || L.cout == 16 || L.cout == 32 || L.cout == 64))) return 1;
  if (L.kind == 1 && P.N % 64 != 0) return 1;

  P.sz = (stride == 2 && L.KD > 1) ? 2 : 1, P.sy = stride, P.sx = stride;
  P.n_phases = P.sz * P.sy * P.sx;
  // per-dimension tap decomposition  k - pad = s*d + r :  halo extent E = dmax - dmin, origin shift o = -dmin
  auto dim_geom = [](int K, int s, int kind, int& E, int& o) {
    if (kind == 1) { E = 1, o = 0; return; }
    const int pad = K / 2;
    if (s == 1) { E = K - 1, o = pad; return; }
    const int dmin = -((pad + 1) / 2), dmax = (K - 1 - pad) / 2;
    E = dmax - dmin, o = -dmin;
  };
  int hz, hy, hx;
  dim_geom(L.KD, P.sz, L.kind, hz, P.oz);
  dim_geom(L.KH, P.sy, L.kind, hy, P.oy);
  dim_geom(L.KH, P.sx, L.kind, hx, P.ox);
  const bool is2d = (L.kind == 0 && L.KD == 1);
  P.TX = 32;
  P.TY = is2d ? 16 : 8;
  P.TZ = is2d ? 1 : 2;
  if (L.kind == 0 && L.KH == 3) {
    if (is2d) P.TY = (L.cin == 8) ? 15 : 7;
    else if (fold) P.TZ = (g_ctas_per_sm >= 2) ? 2 : 4, P.TY = 4;      // sweep: 2 x 4 x 32 with two CTAs per SM (reg0.conv0 42 -> 35 us, reg1.conv0 65 -> 60)
  }
  if (L.kind == 0 && L.KH == 1) P.TY = 8;
  if (L.kind == 1) P.TZ = 2, P.TY = (P.N <= 64) ? 3 : 4;
  // fused lateral: 11 rows (13 x 34 halo pixels, 3 M-tiles = one fold batch) is the tallest tile whose two operand slots + two source
  // slots fit 216 KB; against 7 rows the producer's halo overhead drops 1.29 -> 1.18 and the per-tile fixed costs amortise:
  // 12.7 -> 10.4 ns per pixel in the fused launch (profiles/r2_conv2_timeline.md).  ENERF_B200_LAT_TY overrides (A/B).
  if (lat) {
    static const int lat_ty = [] {
      const char* e = getenv("ENERF_B200_LAT_TY");
      const int v = e ? atoi(e) : 0;
      return (v >= 2 && v <= 15) ? v : 11;
    }();
    P.TY = lat_ty;
  }
  if (stride == 2) {
    if (is2d) P.TY = 7;
    else P.TZ = 2, P.TY = 3;
  }
  if (g_tune2_tz > 0) P.TZ = g_tune2_tz;
  if (g_tune2_ty > 0) P.TY = g_tune2_ty;
  P.TZ = std::min(P.TZ, Dn), P.TY = std::min(P.TY, Hn);

  // the fused-lateral kernel (14-15 warps) runs one CTA per SM whatever the setting
  const bool two_ctas = g_ctas_per_sm >= 2 && !lat;
  const size_t budget = (two_ctas ? 110 : 216) * 1024;
  for (;;) {
    P.IZ = P.TZ + hz, P.IY = P.TY + hy, P.IX = P.TX + hx;
    const int npix = P.IZ * P.IY * P.IX;
    if (P.IZ * P.sz > 256 || P.IY * P.sy > 256 || P.IX * P.sx > 256) return 1;
    const int pmax = ((P.TZ - 1) * P.IY + (P.TY - 1)) * P.IX + P.TX - 1;
    P.n_mt = (pmax + (fold ? 2 : 0)) / 128 + 1;
    P.d128z = 128 / (P.IY * P.IX), P.d128y = (128 % (P.IY * P.IX)) / P.IX, P.d128x = (128 % (P.IY * P.IX)) % P.IX;
    int max_tap_rows = 0, tap_rows[27], tap_phase[27];
    for (int i = 0; i < 27; ++i) tap_phase[i] = 0;
    if (stride == 2) {
      auto split = [](int k, int K, int o, int& d, int& r) {   // k - pad = 2*d + r, returns d' = d + o
        const int tt = k - K / 2;
        const int fl = (tt >= 0) ? tt / 2 : -((-tt + 1) / 2);
        r = tt - 2 * fl, d = fl + o;
      };
      int i = 0;
      for (int kz = 0; kz < L.KD; ++kz)
        for (int ky = 0; ky < L.KH; ++ky)
          for (int kx = 0; kx < L.KH; ++kx) {
            int dz = 0, rz = 0, dy, ry, dx, rx;
            if (P.sz == 2) split(kz, L.KD, P.oz, dz, rz);
            else dz = kz - L.KD / 2 + P.oz;
            split(ky, L.KH, P.oy, dy, ry);
            split(kx, L.KH, P.ox, dx, rx);
            tap_phase[i] = (rz * P.sy + ry) * P.sx + rx;
            tap_rows[i++] = (dz * P.IY + dy) * P.IX + dx;
          }
    } else if (L.kind == 1) {
      for (int d = 0; d < 8; ++d) tap_rows[d] = (((d >> 2) & 1) * P.IY + ((d >> 1) & 1)) * P.IX + (d & 1);
    } else if (fold) {
      for (int kz = 0, i = 0; kz < L.KD; ++kz)
        for (int ky = 0; ky < L.KH; ++ky) tap_rows[i++] = (kz * P.IY + ky) * P.IX;
    } else {
      int i = 0;
      for (int kz = 0; kz < L.KD; ++kz)
        for (int ky = 0; ky < L.KH; ++ky)
          for (int kx = 0; kx < L.KH; ++kx) tap_rows[i++] = (kz * P.IY + ky) * P.IX + kx;
    }
    for (int i = 0; i < n_taps; ++i) max_tap_rows = std::max(max_tap_rows, tap_rows[i]);
    P.n_acc = (2 * P.n_mt * P.N <= (two_ctas ? 256 : 512)) ? 2 : 1;
    const bool tmem_ok = P.n_acc * P.n_mt * P.N <= (two_ctas ? 256 : 512);
    P.xch_bytes = fold ? (uint32_t)align_up((size_t)(P.n_mt * 4 + 1) * 2 * P.N * 4, 64) * 2u : 0u;
    P.FW = P.FH = 0, P.src_slot_bytes = P.src_f1_off = P.src_bytes = 0;
    if (lat) {   // source tiles of the fused lateral: conv0 halo tile (32-byte rows) + the feat1_pre rows / columns its bilinear taps reach
      P.FW = P.IX / 2 + 3, P.FH = P.IY / 2 + 3;
      const uint32_t c0b = (uint32_t)npix * 32u, f1b = (uint32_t)(P.FW * P.FH) * 128u;
      P.src_f1_off = (uint32_t)align_up(c0b, 1024);
      P.src_slot_bytes = P.src_f1_off + (uint32_t)align_up(f1b, 1024);
      P.src_bytes = c0b + f1b;
    }
    // K-block = the widest channel group whose ring still holds >= 2 slots (3-4 preferred)
    bool placed = false;
    const int cands[3] = {32, 16, 8};
    for (int ci = 0; ci < 3 && !placed && tmem_ok; ++ci) {
      const int kbc = lat ? 32 : g_tune2_kbc > 0 ? g_tune2_kbc : cands[ci];
      if (kbc > L.cin || L.cin % kbc) {
        if (g_tune2_kbc > 0) break;
        continue;
      }
      const uint32_t rb = (uint32_t)kbc * 4u;
      // garbage rows of the last M-tile read past the box: keep them inside the slot
      const uint32_t rows_alloc = (uint32_t)std::max(npix, P.n_mt * 128 + max_tap_rows + 2);
      const uint32_t phase_stride = (uint32_t)align_up((size_t)npix * rb, 1024);
      const uint32_t slot = (uint32_t)align_up((size_t)(P.n_phases - 1) * phase_stride + (size_t)rows_alloc * rb, 1024);
      const size_t fixed = 1024 + align_up(P.w_bytes, 128) + P.xch_bytes + 256 + (lat ? 1024 + 2 * (size_t)P.src_slot_bytes : 0);
      int slots = (int)((budget - std::min(budget, fixed)) / slot);
      slots = std::min(slots, 4);
      if (g_tune2_slots > 0) slots = std::min(slots, g_tune2_slots);
      // one CTA per SM: fall to narrower K-blocks rather than a 2-deep ring; two CTAs per SM overlap each other, the widest
      // K-block with two slots wins (sweep: smooth0 59.9 -> 55.8 us, reg1.conv2 20.9 -> 19.4)
      const int want = (ci == 2 || g_tune2_kbc > 0 || two_ctas || lat) ? 2 : 3;
      if (slots >= want || (slots >= 2 && kbc == 8)) {
        P.kbc = kbc, P.n_kb = L.cin / kbc, P.n_slots = slots, P.slot_bytes = slot, P.box_bytes = (uint32_t)npix * rb;
        P.phase_bytes = phase_stride;
        for (int i = 0; i < n_taps; ++i) P.tap_off[i] = (int)((tap_phase[i] * phase_stride + (uint32_t)tap_rows[i] * rb) / 16);
        placed = true;
      }
      if (g_tune2_kbc > 0 || lat) break;
    }
    P.nx = ceil_div(Wn, P.TX), P.ny = ceil_div(Hn, P.TY), P.nz = ceil_div(Dn, P.TZ);
    P.n_tiles = P.nx * P.ny * P.nz;
    // small layers: shrink the tile until every SM has work
    if (placed && (P.n_tiles >= n_sm || (P.TY <= 2 && P.TZ <= 1) || g_tune2_ty > 0)) break;
    if (g_tune2_ty > 0) return 1;
    if (P.TY > 2) P.TY = (P.TY + 1) / 2;
    else if (P.TZ > 1) P.TZ /= 2;
    else if (!placed && P.TX > 16) P.TX /= 2;
    else return 1;
  }
  uint32_t cols = 32;
  while ((int)cols < P.n_acc * P.n_mt * P.N) cols <<= 1;
  P.tmem_cols = cols;
  n_taps_out = n_taps, fold_out = fold;
  return 0;
}

int tc_conv2_try_launch(const TcConvLayer& L, const float* in, int Dn, int Hn, int Wn, const float* wpack, const float* bias, const float* skip,
                        float* out, float* out2, int out_cstride, int out_coff, bool fold_default, cudaStream_t stream, const TcLateral* lat) {
  if (tma::encode_fn() == nullptr) return 1;
  TcConv2Params P;
  int n_taps = 0;
  bool fold = false;
  const int n_sm = device_sm_count();
  if (tc_conv2_plan(L, Dn, Hn, Wn, fold_default, lat != nullptr, bias != nullptr, skip != nullptr, n_sm, P, n_taps, fold, lat ? lat->lat_cin : 8) != 0) return 1;
  P.out_cstride = out_cstride, P.out_coff = out_coff;
  P.wpack = wpack, P.bias = bias, P.skip = skip, P.out = out, P.out2 = out2;
  P.lat_in = lat ? lat->lat_in : nullptr, P.lat_w = lat ? lat->lat_w : nullptr, P.lat_b = lat ? lat->lat_b : nullptr, P.up_in = lat ? lat->up_in : nullptr;
  P.rgb_src = lat ? lat->rgb_src : nullptr, P.packed_out = (lat && lat->rgb_src) ? lat->packed_out : nullptr;
  P.dbg = (g_conv2_dbg_lat_only && lat == nullptr) ? nullptr : g_conv2_dbg;

  // tensor map of the input: channels-last (D,H,W,C) fp32 -> dims {C, W, H, D}, box {kbc, IX, IY, IZ}
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (lat == nullptr) {
    // stride 2: the row grid is the OUTPUT grid, the tensor is the (even) input; a box traverses s*I elements with element stride s
    const uint64_t Wi = (uint64_t)Wn * P.sx, Hi = (uint64_t)Hn * P.sy, Di = (uint64_t)Dn * P.sz;
    const uint64_t dims[4] = {(uint64_t)L.cin, Wi, Hi, Di};
    const uint64_t strides[3] = {(uint64_t)L.cin * 4, Wi * L.cin * 4, Hi * Wi * L.cin * 4};
    const uint32_t box[4] = {(uint32_t)P.kbc, (uint32_t)(P.IX * P.sx), (uint32_t)(P.IY * P.sy), (uint32_t)(P.IZ * P.sz)};
    const uint32_t estr[4] = {1u, (uint32_t)P.sx, (uint32_t)P.sy, (uint32_t)P.sz};
    const int rc = tma::encode_f32(&map, in, 4, dims, strides, box, estr, tma::swizzle_for_bytes(P.kbc * 4));
    ENERF_REQUIRE(rc == 0, ENERF_ECUDA, "tc_conv2: cuTensorMapEncodeTiled failed (%d) for a (%d,%d,%d,%d) tensor, box (%d,%d,%d,%d)", rc, Dn, Hn, Wn, L.cin,
                  P.kbc, P.IX, P.IY, P.IZ);
  }
  CUtensorMap map_c0, map_f1;
  memset(&map_c0, 0, sizeof(map_c0));
  memset(&map_f1, 0, sizeof(map_f1));
  if (lat != nullptr) {
    const uint64_t hi = (uint64_t)Hn / 2, wi = (uint64_t)Wn / 2;
    const uint64_t d0[4] = {8, (uint64_t)Wn, (uint64_t)Hn, (uint64_t)Dn}, s0[3] = {32, (uint64_t)Wn * 32, (uint64_t)Hn * Wn * 32};
    const uint32_t b0[4] = {8, (uint32_t)P.IX, (uint32_t)P.IY, 1};
    int rc = tma::encode_f32(&map_c0, lat->lat_in, 4, d0, s0, b0, nullptr, CU_TENSOR_MAP_SWIZZLE_32B);
    ENERF_REQUIRE(rc == 0, ENERF_ECUDA, "tc_conv2(lateral): cuTensorMapEncodeTiled(conv0) failed (%d)", rc);
    const uint64_t d1[4] = {32, wi, hi, (uint64_t)Dn}, s1[3] = {128, wi * 128, hi * wi * 128};
    const uint32_t b1[4] = {32, (uint32_t)P.FW, (uint32_t)P.FH, 1};
    rc = tma::encode_f32(&map_f1, lat->up_in, 4, d1, s1, b1, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    ENERF_REQUIRE(rc == 0, ENERF_ECUDA, "tc_conv2(lateral): cuTensorMapEncodeTiled(feat1_pre) failed (%d)", rc);
  }
  const size_t smem = 1024 + (size_t)P.n_slots * P.slot_bytes + align_up(P.w_bytes, 128) + P.xch_bytes + 128 + (lat ? 1024 + 2 * (size_t)P.src_slot_bytes : 0);
  const int grid = std::min(P.n_tiles, n_sm * ((g_ctas_per_sm >= 2 && lat == nullptr) ? 2 : 1));
#define TC2_LAUNCH(NT, MD, FD, NM)                                                                                                        \
  do {                                                                                                                                    \
    static PerDeviceSize smem_set_pd;                                                                                                     \
    size_t& smem_set = smem_set_pd.cur();                                                                                                 \
    if (smem > smem_set) {                                                                                                                \
      cudaError_t e = cudaFuncSetAttribute(tc_conv2_kernel<NT, MD, FD, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
      ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tc_conv2: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));               \
      smem_set = smem;                                                                                                                    \
    }                                                                                                                                     \
    tc_conv2_kernel<NT, MD, FD, NM><<<grid, 32 * (4 + NM + 1), smem, stream>>>(map, map_c0, map_f1, P);                                                       \
  } while (0)
// (kx-folded layers always run ONE MMA warp: 6 warps x 2 CTAs = 3 warps per scheduler leave 168 registers per thread for the
//  batched fold epilogue, 7 x 2 would cap it at 128 and spill; two CTAs already give the tensor pipe its two issuing streams)
#define TC2_DISPATCH(NT, MD, FD)             \
  do {                                       \
    if (g_nmma >= 2 && (FD) == 0) TC2_LAUNCH(NT, MD, FD, 2); \
    else TC2_LAUNCH(NT, MD, FD, 1);          \
  } while (0)
  if (lat != nullptr) {
#define TC2_LAT(NT, FD, NM, PR)                                                                                                                  \
  do {                                                                                                                                           \
    static PerDeviceSize lat_set_pd;                                                                                                             \
    size_t& lat_set = lat_set_pd.cur();                                                                                                          \
    if (smem > lat_set) {                                                                                                                        \
      cudaError_t e = cudaFuncSetAttribute(tc_conv2_kernel<NT, TC_PLAIN, FD, NM, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      ENERF_REQUIRE(e == cudaSuccess, ENERF_ECUDA, "tc_conv2(lateral): cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));             \
      lat_set = smem;                                                                                                                            \
    }                                                                                                                                            \
    tc_conv2_kernel<NT, TC_PLAIN, FD, NM, PR><<<grid, 32 * (4 + NM + 1 + PR), smem, stream>>>(map, map_c0, map_f1, P);                                \
  } while (0)
    // (lat1 -> smooth1 is NOT fused: its output feat1_pre is also lat0's up-sampling source, so it must exist in HBM anyway;
    //  the producer is written for lat_cin 8 | 16, only 8 is instantiated)
    // PROD = computing producer warps.  Four (one per scheduler) made the producer the bottleneck of the fused smooth0 (~120 us of
    // latency-bound FP32 work against ~45 us of MMAs); 6 keeps the CTA at 12 warps = 168 registers per thread, 8 halves the
    // producer's time at 128 registers (enerf_tc_conv2_tune's `prod` selects, A/B in profiles/r2_frame_ab.md).
    const int pw = g_prod_warps;
    if (fold) {
      if (pw >= 8) TC2_LAT(3, 1, 1, 8);
      else if (pw >= 6) TC2_LAT(3, 1, 1, 6);
      else TC2_LAT(3, 1, 1, 4);
    } else {
      if (pw >= 8) TC2_LAT(9, 0, 2, 8);
      else if (pw >= 6) TC2_LAT(9, 0, 2, 6);
      else TC2_LAT(9, 0, 2, 4);
    }
#undef TC2_LAT
    ENERF_CHECK_LAUNCH("tc_conv2(lateral)");
    return ENERF_OK;
  }
  if (L.kind == 1) TC2_DISPATCH(8, TC_DECONV, 0);
  else if (fold && n_taps == 9 && L.mode == TC_PLAIN) TC2_DISPATCH(9, TC_PLAIN, 1);
  else if (fold && n_taps == 9 && L.mode == TC_HEAD) TC2_DISPATCH(9, TC_HEAD, 1);
  else if (fold && n_taps == 9 && L.mode == TC_SINGLE) TC2_DISPATCH(9, TC_SINGLE, 1);
  else if (fold && n_taps == 3 && L.mode == TC_PLAIN) TC2_DISPATCH(3, TC_PLAIN, 1);
  else if (!fold && n_taps == 27 && L.mode == TC_PLAIN) TC2_DISPATCH(27, TC_PLAIN, 0);
  else if (!fold && n_taps == 27 && L.mode == TC_HEAD) TC2_DISPATCH(27, TC_HEAD, 0);
  else if (!fold && n_taps == 27 && L.mode == TC_SINGLE) TC2_DISPATCH(27, TC_SINGLE, 0);
  else if (!fold && n_taps == 9 && L.mode == TC_PLAIN) TC2_DISPATCH(9, TC_PLAIN, 0);
  else if (!fold && n_taps == 25 && L.mode == TC_PLAIN) TC2_DISPATCH(25, TC_PLAIN, 0);
  else if (!fold && n_taps == 1 && L.mode == TC_PLAIN) TC2_DISPATCH(1, TC_PLAIN, 0);
  else return 1;
#undef TC2_DISPATCH
#undef TC2_LAUNCH
  ENERF_CHECK_LAUNCH("tc_conv2");
  return ENERF_OK;
}

}  // namespace enerf

// Diagnostic / tuning of the TMA-fed kernel: impl 0 auto | 1 tc_conv.cu only | 2 this kernel where eligible;
// nmma = MMA-issuing warps (1|2); ctas_per_sm (1|2); tz/ty/kbc/slots = forced tile / K-block / ring depth (0 = built-in).
// The launch geometry tc_conv2 would use for a layer on a device with n_sm SMs, without touching a GPU (CPU-testable):
// out[0..39] = TZ TY TX IZ IY IX oz oy ox nx ny nz n_tiles sz sy sx n_phases phase_bytes kbc n_kb n_slots slot_bytes box_bytes
//              N n_mt n_acc n_taps fold tmem_cols w_bytes xch_bytes, out[40..66] = tap_off (16-byte units).  (D,H,W) = row grid.
extern "C" int enerf_tc_conv2_plan(int kind, int KD, int KH, int stride, int cin, int cout, int mode, int D, int H, int W, int fold, int lateral, int n_sm,
                                   int* out) {
  using namespace enerf;
  ENERF_REQUIRE(out && n_sm > 0, ENERF_EINVAL, "tc_conv2_plan: bad arguments");
  TcConvLayer L{kind, KD, KH, cin, cout, mode, 0, stride};
  TcConv2Params P;
  int n_taps = 0;
  bool f = false;
  const int rc = tc_conv2_plan(L, D, H, W, fold != 0, lateral != 0, true, true, n_sm, P, n_taps, f);
  if (rc != 0) {
    set_error("tc_conv2_plan: layer not eligible for the TMA-fed kernel");
    return ENERF_EUNSUPPORTED;
  }
  const int v[40] = {P.TZ, P.TY, P.TX, P.IZ, P.IY, P.IX, P.oz, P.oy, P.ox, P.nx, P.ny, P.nz, P.n_tiles, P.sz, P.sy, P.sx, P.n_phases, (int)P.phase_bytes,
                     P.kbc, P.n_kb, P.n_slots, (int)P.slot_bytes, (int)P.box_bytes, P.N, P.n_mt, P.n_acc, n_taps, f ? 1 : 0, (int)P.tmem_cols,
                     (int)P.w_bytes, (int)P.xch_bytes, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 40; ++i) out[i] = v[i];
  for (int i = 0; i < 27; ++i) out[40 + i] = i < n_taps ? P.tap_off[i] : 0;
  return ENERF_OK;
}

// Diagnostic: when buf != NULL, CTA 0 of every later tc_conv2 launch writes %globaltimer stamps (ns) into buf (device memory,
// 3 roles x 16 tiles x 8 u64): role 0 producer {tile start, first slot free, boxes issued}, 1 MMA warp 0 {start, accumulator
// free, first K-block landed, MMAs issued}, 2 epilogue row 0 {start, accumulator full, accumulator released, outputs stored}.
extern "C" int enerf_tc_conv2_debug(unsigned long long* buf) {
  enerf::g_conv2_dbg = buf;
  enerf::g_conv2_dbg_lat_only = 0;
  return ENERF_OK;
}
// ... only the fused-lateral launch stamps (role 0 = computing producer thread 0: {start, operand slot free, sources landed, tile written})
extern "C" int enerf_tc_conv2_debug_lateral(unsigned long long* buf) {
  enerf::g_conv2_dbg = buf;
  enerf::g_conv2_dbg_lat_only = 1;
  return ENERF_OK;
}

// on: 0 = separate lateral kernel, 1 = fused with the default number of computing producer warps, 4 | 6 | 8 = fused with that many
extern "C" int enerf_tc_conv2_fuse_lateral(int on) {
  ENERF_REQUIRE(on == 0 || on == 1 || on == 4 || on == 6 || on == 8, ENERF_EINVAL, "tc_conv2_fuse_lateral: %d (0 | 1 | 4 | 6 | 8)", on);
  enerf::g_fuse_lateral = on != 0;
  if (on != 0) enerf::g_prod_warps = (on == 1) ? 6 : on;
  return ENERF_OK;
}

extern "C" int enerf_tc_conv2_tune(int impl, int nmma, int ctas_per_sm, int tz, int ty, int kbc, int slots) {
  using namespace enerf;
  ENERF_REQUIRE(impl >= 0 && impl <= 3 && nmma >= 0 && nmma <= 2 && ctas_per_sm >= 0 && ctas_per_sm <= 2 && (kbc == 0 || kbc == 8 || kbc == 16 || kbc == 32),
                ENERF_EINVAL, "tc_conv2_tune: bad arguments");
  if (nmma == 0) nmma = 2;                    // 0 = the shipped default
  if (ctas_per_sm == 0) ctas_per_sm = 2;
  g_impl = impl, g_nmma = nmma, g_ctas_per_sm = ctas_per_sm, g_tune2_tz = tz, g_tune2_ty = ty, g_tune2_kbc = kbc, g_tune2_slots = slots;
  return ENERF_OK;
}
