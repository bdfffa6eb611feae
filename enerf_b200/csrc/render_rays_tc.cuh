// render_rays_tc.cuh -- the fused ray stage with the MLP contractions on the 5th-gen tensor cores.
//
// (kernel template + launcher; instantiated once per view count by render_rays_tc_inst.cu, dispatched by render_rays_tc.cu)
//
// Same scope as render_rays.cuh (build_rays + sample_along_depth + get_vox_feat + get_img_feat +
// Agg + NeRF + raw2outputs; /root/reference/lib/networks/enerf/network.py:24-43), but a CTA owns a
// tile of 128 sample points = the 128 rows (TMEM lanes) of every GEMM:
//
//   thread t gathers point t (trilinear voxel feature, S bilinear image features, ray-diff
//   features), applies view_fc / var / mean on the FP32 pipe and writes its row of the A operands
//   into shared memory (K-major, no swizzle: row r chunk c at c*2048 + r*16 -> conflict-free);
//   one elected thread issues tcgen05.mma kind::tf32 (M=128, accumulators in TMEM):
//     G1  global_fc : per view  [var|mean](24) x 32  +  g_s(16) x 32        (nerf.py:85)
//     G2  fc        : im(32) x 16                                            (nerf.py:88)
//     G3  lr0       : vox_img_feat(24) x 64                                  (nerf.py:33)
//     G5  color.0   : per view  [x|vif](88) x 64  +  [f_s|dir_s](16) x 64    (nerf.py:38-40)
//   and the 128 threads run the epilogues out of TMEM (tcgen05.ld): bias+ReLU, the 32->1 / 64->1
//   dot products (agg_w_fc, sigma, color.2), the two softmaxes over views, the rgb blend, and the
//   alpha compositing over the samples of a ray with warp shuffles (samples of a ray sit in
//   adjacent lanes).  tcgen05.commit -> mbarrier hands each accumulator to the epilogue.
//
// Numerics: GEMM operands are rounded to TF32 (cvt.rna), accumulation and everything else fp32.
// Measured effect of TF32 MLP operands on the reference itself: max|d rgb| 2.8e-4, dPSNR 7e-5 dB
//  -- 100x inside the |dPSNR| < 0.01 dB budget.  The FP32-pipe kernel
// (render_rays.cuh) stays available as the exact mode.
#pragma once
#include "common.cuh"
#include "render_rays_params.cuh"
#include "tc.cuh"

namespace enerf {

constexpr int TC_FC = 11;        // image feature (8) + rgb (3)
constexpr int TC_FCP = 12;

// float offsets inside the packed weight blob (host: enerf_b200/packing.py::pack_nerf_tc)
struct TcW {
  static constexpr int bg_view = 0;                      // B [4 chunks][32][4]   K=16 N=32
  static constexpr int bg_shared = bg_view + 4 * 32 * 4; // B [6][32][4]          K=24 N=32
  static constexpr int bfc = bg_shared + 6 * 32 * 4;     // B [8][16][4]          K=32 N=16
  static constexpr int b0 = bfc + 8 * 16 * 4;            // B [6][64][4]          K=24 N=64
  static constexpr int bc_shared = b0 + 6 * 64 * 4;      // B [22][64][4]         K=88 N=64
  static constexpr int bc_view = bc_shared + 22 * 64 * 4;  // B [4][64][4]        K=16 N=64
  static constexpr int v_view_w = bc_view + 4 * 64 * 4;  // [4][12]
  static constexpr int v_view_b = v_view_w + 48;         // [12]
  static constexpr int v_bg = v_view_b + 12;             // [32]
  static constexpr int v_wa = v_bg + 32;                 // [32]
  static constexpr int v_ba = v_wa + 32;                 // [4]
  static constexpr int v_bf = v_ba + 4;                  // [16]
  static constexpr int v_b0 = v_bf + 16;                 // [64]
  static constexpr int v_ws = v_b0 + 64;                 // [64]
  static constexpr int v_bs = v_ws + 64;                 // [4]
  static constexpr int v_bc = v_bs + 4;                  // [64]
  static constexpr int v_w2 = v_bc + 64;                 // [64]
  static constexpr int v_b2 = v_w2 + 64;                 // [4]
  static constexpr int total = v_b2 + 4;                 // 10,392 floats
};

struct RayTcParams {
  RayParams r;
  const float* wblob;
  int n_tiles;
  unsigned long long* dbg;   // optional phase stamps (ns) of CTA 0's first two tiles
};

__device__ __forceinline__ unsigned long long ray_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define RAY_STAMP(i_)                                                          \
  do {                                                                         \
    if (dbg && tile_i < 2) dbg[tile_i * 16 + (i_)] = ray_gtime();              \
  } while (0)

template <int S>
struct TcSmem {
  // region P: [x (16) | vox (2) | img (4)] = 22 chunks; before G1 it holds g_s (4 per view) and [var | mean] (6), so it
  // must span 4S + 6 chunks when that is more (5-8 views); the per-view [f | dir] chunks follow it
  static constexpr int P_CHUNKS = (4 * S + 6 > 22) ? 4 * S + 6 : 22;
  static constexpr int A_CHUNKS = P_CHUNKS + 4 * S;
  static constexpr int CHUNK_FLOATS = 128 * 4;           // 2 KB
  static constexpr int o_w = 0;
  static constexpr int o_a = (TcW::total + 31) / 32 * 32;
  static constexpr int o_cam = o_a + A_CHUNKS * CHUNK_FLOATS;
  static constexpr int total_floats = o_cam + ENERF_MAX_VIEWS * 24 + 4;
  static constexpr size_t bytes = (size_t)total_floats * 4 + 64;
};

// The TF32-mode kernel's scalar math: approximate division / square root / exponential (MUFU + one multiply instead of the
// IEEE sequences: ~12 instructions and a slow-path branch per division, ~50 for log1pf(expf)).  ncu counted 51 divisions per
// sample point = ~13 % of the kernel's issued instructions, all on dependent chains of a latency-bound kernel.  Errors are a
// few ulp -- two orders below the TF32 operand rounding this kernel already has (the FP32-pipe kernel render_rays.cuh stays exact).
#ifndef RTC_FAST_MATH
#define RTC_FAST_MATH 1
#endif
// (.ftz forms: the non-ftz approximations carry a range-fixup compare + scale around every MUFU)
__device__ __forceinline__ float mufu_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rdiv(float a, float b) {
#if RTC_FAST_MATH
  return a * mufu_rcp(b);
#else
  return a / b;
#endif
}
__device__ __forceinline__ float rsqrt_(float x) {
#if RTC_FAST_MATH
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
__device__ __forceinline__ float rexp(float x) {
#if RTC_FAST_MATH
  return mufu_ex2(x * 1.4426950408889634f);
#else
  return expf(x);
#endif
}
__device__ __forceinline__ float rsoftplus(float x) {      // log(1 + exp(x)), x <= 20
#if RTC_FAST_MATH
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + mufu_ex2(x * 1.4426950408889634f)));
  return r * 0.6931471805599453f;
#else
  return log1pf(expf(x));
#endif
}

__device__ __forceinline__ void store_chunk(float* a_base, int chunk, int row, float v0, float v1, float v2, float v3) {
  *reinterpret_cast<float4*>(a_base + (size_t)chunk * 512 + row * 4) =
      make_float4(tc::to_tf32(v0), tc::to_tf32(v1), tc::to_tf32(v2), tc::to_tf32(v3));
}

template <int S>
__global__ void __launch_bounds__(128, (S <= 4) ? 2 : 1) render_rays_tc_kernel(const RayTcParams P) {
  using SM = TcSmem<S>;
  extern __shared__ __align__(128) float smem[];
  float* sw = smem + SM::o_w;
  float* a_s = smem + SM::o_a;
  float* cam_s = smem + SM::o_cam;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const RayParams& p = P.r;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // color.0 FACTORED (S <= 3, where the extra accumulator still fits 2 x 256 TMEM columns per SM): its view-shared K (x | vox | img: 88 of
  // the 104 columns) is multiplied ONCE into accumulator 0 and only the per-view [f | dir] K = 16 per view -- 11 + 2 S MMAs instead of 13 S
  // (17 vs 39 at S = 3); the epilogue adds the two accumulators before the ReLU.
  constexpr bool FACT = (S <= 3);
  constexpr int G5_ACCS = FACT ? S + 1 : S;
  constexpr uint32_t TMEM_COLS = (G5_ACCS * 64 <= 128) ? 128 : (G5_ACCS * 64 <= 256) ? 256 : 512;   // G5 holds 64-column accumulators

  // ---- one-time setup: weights, cameras, barrier, TMEM ----
  for (int e = t; e < TcW::total; e += 128) sw[e] = __ldg(P.wblob + e);
  {
    const EnerfCam* cam = p.cam;
    for (int e = t; e < S * 24; e += 128) {
      const int s = e / 24, k = e % 24;
      cam_s[e] = (k < 12) ? cam->src_ext[s][k] : (k < 21) ? cam->src_ixt[p.level][s][k - 12] : cam->src_center[s][k - 21];
    }
    if (t < 3) cam_s[ENERF_MAX_VIEWS * 24 + t] = cam->tar_center[t];
  }
  if (t == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc::fence_proxy_async();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);     // warp-uniform copy for the MMA issue
  const uint32_t a_addr = tc::smem_u32(a_s), w_addr = tc::smem_u32(sw);
  uint32_t phase = 0;

  const int Ns = p.num_samples;
  const int ns_shift = 31 - __clz(Ns);
  const float tcx = cam_s[ENERF_MAX_VIEWS * 24 + 0], tcy = cam_s[ENERF_MAX_VIEWS * 24 + 1], tcz = cam_s[ENERF_MAX_VIEWS * 24 + 2];
  const size_t hw = (size_t)p.hv * p.wv;

  auto a_desc = [&](int chunk) { return tc::smem_desc(a_addr + (uint32_t)chunk * 2048u, 2048u, 128u); };
  auto b_desc = [&](int off_floats, int chunk, int N) {
    return tc::smem_desc(w_addr + (uint32_t)off_floats * 4u + (uint32_t)chunk * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
  };
  auto sync_then_issue = [&]() {  // A rows written by everyone -> visible to the tensor core
    tc::fence_proxy_async();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
  };
  auto wait_mma = [&]() {
    tc::mbar_wait(&bar, phase);
    phase ^= 1;
    tc::tc_fence_after_sync();
  };

  unsigned long long* dbg = (P.dbg && t == 0 && blockIdx.x == 0) ? P.dbg : nullptr;
  int tile_i = -1;
  // device-side ray count (masked rays without a host read-back): the grid was sized for the upper bound p.n_rays
  const int n_rays = p.n_rays_dev ? min(__ldg(p.n_rays_dev), p.n_rays) : p.n_rays;
  const int n_tiles = p.n_rays_dev ? (int)(((long long)n_rays * Ns + 127) / 128) : P.n_tiles;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    ++tile_i;
    RAY_STAMP(0);
    const int pt = tile * 128 + t;                        // < 2^31: the launcher checks n_rays * Ns
    const bool valid = pt < n_rays * Ns;
    int ray = valid ? (pt >> ns_shift) : n_rays - 1;      // Ns is 1 | 2 | 4 | 8 on this path
    if (p.out_raw) ray = (p.win_y + ray / p.win_w) * p.Wr + p.win_x + ray % p.win_w;   // layered mode: window -> frame pixel
    const int k = pt & (Ns - 1);

    // ================= stage A: build_rays + sample + gathers (FP32 pipe) =================
    const float4 r0 = ldg4(p.rays + (size_t)ray * 8), r1 = ldg4(p.rays + (size_t)ray * 8 + 4);
    const float u = r1.z, v = r1.w;
    const int ui = (int)u, vi = (int)v;
    float dep, sd, vn, vf;
    if (p.hv == p.Hr && p.wv == p.Wr) {
      const size_t o = (size_t)vi * p.wv + ui;
      dep = __ldg(p.depth + o), sd = __ldg(p.std + o), vn = __ldg(p.near_far + o), vf = __ldg(p.near_far + hw + o);
    } else {
      dep = bilinear_ac(p.depth, p.hv, p.wv, p.Hr, p.Wr, vi, ui);
      sd = bilinear_ac(p.std, p.hv, p.wv, p.Hr, p.Wr, vi, ui);
      vn = bilinear_ac(p.near_far, p.hv, p.wv, p.Hr, p.Wr, vi, ui);
      vf = bilinear_ac(p.near_far + hw, p.hv, p.wv, p.Hr, p.Wr, vi, ui);
    }
    float rn, rf;
    if (p.depth_inv) {
      rn = fminf(dep + sd, vn);
      rf = fmaxf(dep - sd, vf);
    } else {
      rn = fmaxf(dep - sd, vn);
      rf = fminf(dep + sd, vf);
    }
    const float z = (Ns == 1) ? rn + (rf - rn) * 0.5f : rn + (rf - rn) * linspace01(k, Ns);
    const float tz = p.depth_inv ? rdiv(1.0f, fmaxf(z, 1e-6f)) : z;
    const float X = r0.x + r0.w * tz, Y = r0.y + r1.x * tz, Z = r0.z + r1.y * tz;
    const float dn = p.depth_inv ? rdiv(vn - z, fmaxf(vn - vf, 1e-6f)) : rdiv(z - vn, fmaxf(vf - vn, 1e-6f));

    float vox[8];
    {
#if RTC_FAST_MATH
      const float ix = rdiv(u, (float)(p.Wr - 1)) * (float)(p.wv - 1), iy = rdiv(v, (float)(p.Hr - 1)) * (float)(p.hv - 1), iz = dn * (float)(p.D - 1);
#else
      const float un = u / (float)(p.Wr - 1), vnrm = v / (float)(p.Hr - 1);
      const float gx = un * 2.f - 1.f, gy = vnrm * 2.f - 1.f, gz = dn * 2.f - 1.f;
      const float ix = ((gx + 1.f) / 2.f) * (float)(p.wv - 1), iy = ((gy + 1.f) / 2.f) * (float)(p.hv - 1),
                  iz = ((gz + 1.f) / 2.f) * (float)(p.D - 1);
#endif
#pragma unroll
      for (int c = 0; c < 8; ++c) vox[c] = 0.f;
      if (p.feat_vol != nullptr && ix > -1.f && ix < (float)p.wv && iy > -1.f && iy < (float)p.hv && iz > -1.f && iz < (float)p.D) {
        const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
        const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
        const float wx[2] = {(fx0 + 1.f) - ix, ix - fx0}, wy[2] = {(fy0 + 1.f) - iy, iy - fy0}, wz[2] = {(fz0 + 1.f) - iz, iz - fz0};
#pragma unroll
        for (int cz = 0; cz < 2; ++cz)
#pragma unroll
          for (int cy = 0; cy < 2; ++cy)
#pragma unroll
            for (int cx = 0; cx < 2; ++cx) {
              const int xx = x0 + cx, yy = y0 + cy, zz = z0 + cz;
              if (xx >= 0 && xx < p.wv && yy >= 0 && yy < p.hv && zz >= 0 && zz < p.D && yy >= p.vol_y0 && yy < p.vol_y0 + p.vol_h) {
                const float wgt = wx[cx] * wy[cy] * wz[cz];
                const float* q = p.feat_vol + (((size_t)zz * p.vol_h + (yy - p.vol_y0)) * p.wv + xx) * 8;
                const float4 a = ldg4(q), b = ldg4(q + 4);
                vox[0] = fmaf(a.x, wgt, vox[0]), vox[1] = fmaf(a.y, wgt, vox[1]), vox[2] = fmaf(a.z, wgt, vox[2]),
                vox[3] = fmaf(a.w, wgt, vox[3]);
                vox[4] = fmaf(b.x, wgt, vox[4]), vox[5] = fmaf(b.y, wgt, vox[5]), vox[6] = fmaf(b.z, wgt, vox[6]),
                vox[7] = fmaf(b.w, wgt, vox[7]);
              }
            }
      }
    }

    float g[S][TC_FCP];      // image feature + rgb, later + relu(view_fc(dir))
    float rgb_s[S][3];
    float tdx = X - tcx, tdy = Y - tcy, tdz = Z - tcz;
    {
      const float n = rsqrt_(tdx * tdx + tdy * tdy + tdz * tdz) + 1e-6f;
      tdx = rdiv(tdx, n), tdy = rdiv(tdy, n), tdz = rdiv(tdz, n);
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float* E = cam_s + s * 24;
      const float* K = E + 12;
      const float* Cn = E + 21;
      const float xc = E[0] * X + E[1] * Y + E[2] * Z + E[3];
      const float yc = E[4] * X + E[5] * Y + E[6] * Z + E[7];
      const float zc = E[8] * X + E[9] * Y + E[10] * Z + E[11];
      const float p0 = K[0] * xc + K[1] * yc + K[2] * zc;
      const float p1 = K[3] * xc + K[4] * yc + K[5] * zc;
      const float p2 = K[6] * xc + K[7] * yc + K[8] * zc;
      const float pz = fmaxf(p2, 1e-6f);
#if RTC_FAST_MATH
      // the reference normalises to [-1,1] and grid_sample(align_corners=True) maps straight back: the pixel coordinate itself
      float ix = rdiv(p0, pz), iy = rdiv(p1, pz);
#else
      const float gx = (p0 / pz) / (float)(p.Wr - 1) * 2.f - 1.f, gy = (p1 / pz) / (float)(p.Hr - 1) * 2.f - 1.f;
      float ix = ((gx + 1.f) / 2.f) * (float)(p.Wr - 1), iy = ((gy + 1.f) / 2.f) * (float)(p.Hr - 1);
#endif
      ix = fminf(fmaxf(ix, 0.f), (float)(p.Wr - 1));
      iy = fminf(fmaxf(iy, 0.f), (float)(p.Hr - 1));
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      const int x0 = (int)fx0, y0 = (int)fy0;
      const int x1 = min(x0 + 1, p.Wr - 1), y1 = min(y0 + 1, p.Hr - 1);
      const float txr = (fx0 + 1.f) - ix, txl = ix - fx0, tyb = (fy0 + 1.f) - iy, tyt = iy - fy0;
      const float w_nw = txr * tyb, w_ne = txl * tyb, w_sw = txr * tyt, w_se = txl * tyt;
      const float* base = p.img + (size_t)s * p.Hr * p.Wr * TC_FCP;
      const float* q00 = base + ((size_t)y0 * p.Wr + x0) * TC_FCP;
      const float* q01 = base + ((size_t)y0 * p.Wr + x1) * TC_FCP;
      const float* q10 = base + ((size_t)y1 * p.Wr + x0) * TC_FCP;
      const float* q11 = base + ((size_t)y1 * p.Wr + x1) * TC_FCP;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 a = ldg4(q00 + 4 * q), b = ldg4(q01 + 4 * q), c = ldg4(q10 + 4 * q), d = ldg4(q11 + 4 * q);
        g[s][4 * q + 0] = a.x * w_nw + b.x * w_ne + c.x * w_sw + d.x * w_se;
        g[s][4 * q + 1] = a.y * w_nw + b.y * w_ne + c.y * w_sw + d.y * w_se;
        g[s][4 * q + 2] = a.z * w_nw + b.z * w_ne + c.z * w_sw + d.z * w_se;
        g[s][4 * q + 3] = a.w * w_nw + b.w * w_ne + c.w * w_sw + d.w * w_se;
      }
      rgb_s[s][0] = g[s][8], rgb_s[s][1] = g[s][9], rgb_s[s][2] = g[s][10];
      float sx = X - Cn[0], sy = Y - Cn[1], sz = Z - Cn[2];
      const float n = rsqrt_(sx * sx + sy * sy + sz * sz) + 1e-6f;
      sx = rdiv(sx, n), sy = rdiv(sy, n), sz = rdiv(sz, n);
      const float rx = tdx - sx, ry = tdy - sy, rz = tdz - sz;
      const float rnm = fmaxf(rsqrt_(rx * rx + ry * ry + rz * rz), 1e-6f);
      const float d0 = rdiv(rx, rnm), d1 = rdiv(ry, rnm), d2 = rdiv(rz, rnm), d3 = tdx * sx + tdy * sy + tdz * sz;
      // A operand of color.0's per-view columns: [f_s (11) | dir_s (4) | 0]
      const int cf = SM::P_CHUNKS + 4 * s;
      store_chunk(a_s, cf + 0, t, g[s][0], g[s][1], g[s][2], g[s][3]);
      store_chunk(a_s, cf + 1, t, g[s][4], g[s][5], g[s][6], g[s][7]);
      store_chunk(a_s, cf + 2, t, g[s][8], g[s][9], g[s][10], d0);
      store_chunk(a_s, cf + 3, t, d1, d2, d3, 1.f);   // column 15 = 1: color.0's bias rides in B (row 15)
      // g_s = f_s + relu(view_fc(dir_s))    (nerf.py:76-78)
      if (p.viewdir_agg) {
#pragma unroll
        for (int c = 0; c < TC_FC; ++c) {
          float a = sw[TcW::v_view_b + c];
          a = fmaf(d0, sw[TcW::v_view_w + 0 * 12 + c], a);
          a = fmaf(d1, sw[TcW::v_view_w + 1 * 12 + c], a);
          a = fmaf(d2, sw[TcW::v_view_w + 2 * 12 + c], a);
          a = fmaf(d3, sw[TcW::v_view_w + 3 * 12 + c], a);
          g[s][c] += fmaxf(a, 0.f);
        }
      }
      store_chunk(a_s, 4 * s + 0, t, g[s][0], g[s][1], g[s][2], g[s][3]);
      store_chunk(a_s, 4 * s + 1, t, g[s][4], g[s][5], g[s][6], g[s][7]);
      store_chunk(a_s, 4 * s + 2, t, g[s][8], g[s][9], g[s][10], 1.f);   // column 11 = 1: global_fc's bias rides in B
      store_chunk(a_s, 4 * s + 3, t, 0.f, 0.f, 0.f, 0.f);
    }
    {
      float vr[12], mn[12];
      vr[11] = 0.f, mn[11] = 0.f;
      const float invS = 1.0f / (float)S, invS1 = 1.0f / (float)(S - 1);
#pragma unroll
      for (int c = 0; c < TC_FC; ++c) {
        float m = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) m += g[s][c];
        m *= invS;
        float q = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) q += (g[s][c] - m) * (g[s][c] - m);
        vr[c] = q * invS1, mn[c] = m;
      }
      const int cv = 4 * S;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        store_chunk(a_s, cv + q, t, vr[4 * q], vr[4 * q + 1], vr[4 * q + 2], vr[4 * q + 3]);
        store_chunk(a_s, cv + 3 + q, t, mn[4 * q], mn[4 * q + 1], mn[4 * q + 2], mn[4 * q + 3]);
      }
    }

    RAY_STAMP(1);   // gathers + A rows written
    // ================= G1: global_fc =================
    sync_then_issue();
    RAY_STAMP(2);
    if (warp == 0) {   // converged warp, uniform operands, one elected lane issues (tc::mma_tf32_elect)
     {
      const uint32_t id = tc::idesc_tf32(128, 32);
#pragma unroll
      for (int s = 0; s < S; ++s) {
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) tc::mma_tf32_elect(tmem_u + s * 32, a_desc(4 * S + 2 * kk), b_desc(TcW::bg_shared, 2 * kk, 32), id, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) tc::mma_tf32_elect(tmem_u + s * 32, a_desc(4 * s + 2 * kk), b_desc(TcW::bg_view, 2 * kk, 32), id, 1);
      }
      tc::mma_commit_elect(&bar);
     }
     __syncwarp();
    }
    wait_mma();
    RAY_STAMP(3);
    // ---- E1: ReLU, agg_w_fc logits, softmax over views, weighted pooling (one pass over TMEM:
    //      running max + rescale, mathematically the max-subtracted softmax of nerf.py:87) ----
    {
      float im[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) im[j] = 0.f;
      float mx = -INFINITY, den = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        float h[32];
        tc::tmem_ld32(tmem_row + s * 32, h);
        tc::tmem_ld_wait();
        // 32 -> 1 dot product as packed f32x2 FMAs on two independent chains (a 32-deep dependent FFMA chain is pure latency
        // in a kernel with two warps per scheduler)
        float2 d0 = make_float2(sw[TcW::v_ba], 0.f), d1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          h[j] = fmaxf(h[j], 0.f), h[j + 1] = fmaxf(h[j + 1], 0.f), h[j + 2] = fmaxf(h[j + 2], 0.f), h[j + 3] = fmaxf(h[j + 3], 0.f);   // bias already in the accumulator
          const float4 w4 = *reinterpret_cast<const float4*>(sw + TcW::v_wa + j);
          d0 = __ffma2_rn(make_float2(h[j], h[j + 1]), make_float2(w4.x, w4.y), d0);
          d1 = __ffma2_rn(make_float2(h[j + 2], h[j + 3]), make_float2(w4.z, w4.w), d1);
        }
        const float a = (d0.x + d0.y) + (d1.x + d1.y);
        const float lg = fmaxf(a, 0.f);
        const float mn = fmaxf(mx, lg);
        const float sc = rexp(mx - mn), e = rexp(lg - mn);   // first view: sc = exp(-inf) = 0
        den = fmaf(den, sc, e);
        const float2 e2 = make_float2(e, e), sc2 = make_float2(sc, sc);
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 r = __ffma2_rn(e2, make_float2(h[j], h[j + 1]), __fmul2_rn(make_float2(im[j], im[j + 1]), sc2));
          im[j] = r.x, im[j + 1] = r.y;
        }
        mx = mn;
      }
      const float inv = rdiv(1.0f, den);
#pragma unroll
      for (int q = 0; q < 8; ++q) store_chunk(a_s, q, t, im[4 * q] * inv, im[4 * q + 1] * inv, im[4 * q + 2] * inv, im[4 * q + 3] * inv);
    }
    RAY_STAMP(4);
    // ================= G2: fc 32 -> 16 =================
    sync_then_issue();
    if (warp == 0) {   // converged warp, uniform operands, one elected lane issues (tc::mma_tf32_elect)
     {
      const uint32_t id = tc::idesc_tf32(128, 16);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) tc::mma_tf32_elect(tmem_u, a_desc(2 * kk), b_desc(TcW::bfc, 2 * kk, 16), id, kk > 0);
      tc::mma_commit_elect(&bar);
     }
     __syncwarp();
    }
    wait_mma();
    RAY_STAMP(5);
    {
      float o[16];
      tc::tmem_ld16(tmem_row, o);
      tc::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) o[c] = fmaxf(o[c] + sw[TcW::v_bf + c], 0.f);
      // vox_img_feat = [vox(8) | img(16)] -> chunks 16..21 (A of lr0 and of color.0's shared part)
      store_chunk(a_s, 16, t, vox[0], vox[1], vox[2], vox[3]);
      store_chunk(a_s, 17, t, vox[4], vox[5], vox[6], vox[7]);
#pragma unroll
      for (int q = 0; q < 4; ++q) store_chunk(a_s, 18 + q, t, o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    }
    RAY_STAMP(6);
    // ================= G3: lr0 24 -> 64 =================
    sync_then_issue();
    if (warp == 0) {   // converged warp, uniform operands, one elected lane issues (tc::mma_tf32_elect)
     {
      const uint32_t id = tc::idesc_tf32(128, 64);
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) tc::mma_tf32_elect(tmem_u, a_desc(16 + 2 * kk), b_desc(TcW::b0, 2 * kk, 64), id, kk > 0);
      tc::mma_commit_elect(&bar);
     }
     __syncwarp();
    }
    wait_mma();
    RAY_STAMP(7);
    float sigma;
    {
      float2 g0 = make_float2(sw[TcW::v_bs], 0.f), g1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float x[32];
        tc::tmem_ld32(tmem_row + half * 32, x);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sw + TcW::v_b0 + half * 32 + j);
          const float4 w4 = *reinterpret_cast<const float4*>(sw + TcW::v_ws + half * 32 + j);
          const float2 s0 = __fadd2_rn(make_float2(x[j], x[j + 1]), make_float2(b4.x, b4.y));
          const float2 s1 = __fadd2_rn(make_float2(x[j + 2], x[j + 3]), make_float2(b4.z, b4.w));
          x[j] = fmaxf(s0.x, 0.f), x[j + 1] = fmaxf(s0.y, 0.f), x[j + 2] = fmaxf(s1.x, 0.f), x[j + 3] = fmaxf(s1.y, 0.f);
          g0 = __ffma2_rn(make_float2(x[j], x[j + 1]), make_float2(w4.x, w4.y), g0);
          g1 = __ffma2_rn(make_float2(x[j + 2], x[j + 3]), make_float2(w4.z, w4.w), g1);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) store_chunk(a_s, half * 8 + q, t, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
      }
      const float sg = (g0.x + g0.y) + (g1.x + g1.y);
      sigma = (sg > 20.f) ? sg : rsoftplus(sg);
    }
    RAY_STAMP(8);
    // ================= G5: color.0 =================
    sync_then_issue();
    RAY_STAMP(9);
    if (warp == 0) {   // converged warp, uniform operands, one elected lane issues (tc::mma_tf32_elect)
     {
      const uint32_t id = tc::idesc_tf32(128, 64);
#pragma unroll
      if constexpr (FACT) {
#pragma unroll
        for (int kk = 0; kk < 11; ++kk) tc::mma_tf32_elect(tmem_u, a_desc(2 * kk), b_desc(TcW::bc_shared, 2 * kk, 64), id, kk > 0);
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            tc::mma_tf32_elect(tmem_u + 64 + s * 64, a_desc(SM::P_CHUNKS + 4 * s + 2 * kk), b_desc(TcW::bc_view, 2 * kk, 64), id, kk > 0);
        }
      } else {
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int kk = 0; kk < 11; ++kk) tc::mma_tf32_elect(tmem_u + s * 64, a_desc(2 * kk), b_desc(TcW::bc_shared, 2 * kk, 64), id, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) tc::mma_tf32_elect(tmem_u + s * 64, a_desc(SM::P_CHUNKS + 4 * s + 2 * kk), b_desc(TcW::bc_view, 2 * kk, 64), id, 1);
        }
      }
      tc::mma_commit_elect(&bar);
     }
     __syncwarp();
    }
    wait_mma();
    RAY_STAMP(10);
    float cr = 0.f, cg = 0.f, cb = 0.f;
    {
      float cl[S];
      float mx = -INFINITY;
      if constexpr (FACT) {
        float2 c0[S], c1[S];
#pragma unroll
        for (int s = 0; s < S; ++s) c0[s] = make_float2(sw[TcW::v_b2], 0.f), c1[s] = make_float2(0.f, 0.f);
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {           // 16 of the 64 columns at a time (registers: the kernel stays under 200)
          float sh[16];                            // the view-shared product (accumulator 0)
          tc::tmem_ld16(tmem_row + qt * 16, sh);
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float h[16];
            tc::tmem_ld16(tmem_row + 64 + s * 64 + qt * 16, h);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {       // bias in the per-view accumulator; packed adds / FMAs on two chains
              const float4 w4 = *reinterpret_cast<const float4*>(sw + TcW::v_w2 + qt * 16 + j);
              const float2 a0 = __fadd2_rn(make_float2(h[j], h[j + 1]), make_float2(sh[j], sh[j + 1]));
              const float2 a1 = __fadd2_rn(make_float2(h[j + 2], h[j + 3]), make_float2(sh[j + 2], sh[j + 3]));
              c0[s] = __ffma2_rn(make_float2(fmaxf(a0.x, 0.f), fmaxf(a0.y, 0.f)), make_float2(w4.x, w4.y), c0[s]);
              c1[s] = __ffma2_rn(make_float2(fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f)), make_float2(w4.z, w4.w), c1[s]);
            }
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float a = (c0[s].x + c0[s].y) + (c1[s].x + c1[s].y);
          cl[s] = fmaxf(a, 0.f);
          mx = fmaxf(mx, cl[s]);
        }
      } else {
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float2 c0 = make_float2(sw[TcW::v_b2], 0.f), c1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float h[32];
            tc::tmem_ld32(tmem_row + s * 64 + half * 32, h);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {       // bias in the accumulator; packed FMAs on two chains
              const float4 w4 = *reinterpret_cast<const float4*>(sw + TcW::v_w2 + half * 32 + j);
              c0 = __ffma2_rn(make_float2(fmaxf(h[j], 0.f), fmaxf(h[j + 1], 0.f)), make_float2(w4.x, w4.y), c0);
              c1 = __ffma2_rn(make_float2(fmaxf(h[j + 2], 0.f), fmaxf(h[j + 3], 0.f)), make_float2(w4.z, w4.w), c1);
            }
          }
          const float a = (c0.x + c0.y) + (c1.x + c1.y);
          cl[s] = fmaxf(a, 0.f);
          mx = fmaxf(mx, cl[s]);
        }
      }
      float den = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        cl[s] = rexp(cl[s] - mx);
        den += cl[s];
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float ws_ = rdiv(cl[s], den);
        cr = fmaf(rgb_s[s][0], ws_, cr), cg = fmaf(rgb_s[s][1], ws_, cg), cb = fmaf(rgb_s[s][2], ws_, cb);
      }
    }

    RAY_STAMP(11);
    // ================= raw2outputs: prefix product / sums over the Ns lanes of a ray =================
    if (p.out_raw) {   // layered mode: samples are merged across layers by enerf_composite_layers
      if (valid) {
        const size_t o = (size_t)ray * p.out_stride + p.out_off + k;
        *reinterpret_cast<float4*>(p.out_raw + o * 4) = make_float4(cr, cg, cb, sigma);
        p.out_z[o] = p.depth_inv ? rdiv(1.0f, z) : z;
      }
    } else if (Ns == 2) {
      // two samples of a ray in lanes (2i, 2i+1): one xor-shuffle exchange per quantity
      const float alpha = 1.f - rexp(-sigma);
      const float tr = 1.f - alpha + 1e-10f;
      const float tr_o = __shfl_xor_sync(0xffffffffu, tr, 1);
      const float wk = alpha * (k == 0 ? 1.f : tr_o);
      const float wk_o = __shfl_xor_sync(0xffffffffu, wk, 1);
      const float cr_o = __shfl_xor_sync(0xffffffffu, cr, 1), cg_o = __shfl_xor_sync(0xffffffffu, cg, 1),
                  cb_o = __shfl_xor_sync(0xffffffffu, cb, 1), z_o = __shfl_xor_sync(0xffffffffu, z, 1);
      const float mx = fmaxf(wk, wk_o);
      const float e = rexp(wk - mx), e_o = rexp(wk_o - mx);
      if (valid) {
        if (k == 0) {
          const float den = e + e_o;                       // same order as the sequential sum over samples
          const float wn0 = rdiv(e, den), wn1 = rdiv(e_o, den);
          float ar = fmaf(wk_o, cr_o, wk * cr), ag = fmaf(wk_o, cg_o, wk * cg), ab = fmaf(wk_o, cb_o, wk * cb);
          if (p.white_bkgd) {
            const float bg = 1.f - (wn0 + wn1);
            ar += bg, ag += bg, ab += bg;
          }
          *reinterpret_cast<float2*>(p.out_weights + (size_t)ray * 2) = make_float2(wn0, wn1);
          p.out_rgb[(size_t)ray * 3 + 0] = ar;
          p.out_rgb[(size_t)ray * 3 + 1] = ag;
          p.out_rgb[(size_t)ray * 3 + 2] = ab;
          p.out_depth[ray] = wn0 * z + wn1 * z_o;
        }
      }
    } else {
      const int gbase = lane - k;                 // first lane of this ray's group (Ns | 32)
      const float alpha = 1.f - rexp(-sigma);
      const float tr = 1.f - alpha + 1e-10f;
      float T = 1.f;
      for (int j = 0; j + 1 < Ns; ++j) {
        const float tj = __shfl_sync(0xffffffffu, tr, gbase + j);
        if (j < k) T *= tj;
      }
      const float wk = alpha * T;
      float ar = 0.f, ag = 0.f, ab = 0.f, mx = -INFINITY;
      for (int j = 0; j < Ns; ++j) {
        const float wj = __shfl_sync(0xffffffffu, wk, gbase + j);
        ar = fmaf(wj, __shfl_sync(0xffffffffu, cr, gbase + j), ar);
        ag = fmaf(wj, __shfl_sync(0xffffffffu, cg, gbase + j), ag);
        ab = fmaf(wj, __shfl_sync(0xffffffffu, cb, gbase + j), ab);
        mx = fmaxf(mx, wj);
      }
      const float e = rexp(wk - mx);
      float den = 0.f;
      for (int j = 0; j < Ns; ++j) den += __shfl_sync(0xffffffffu, e, gbase + j);
      const float wn = rdiv(e, den);
      float dsum = 0.f, wsum = 0.f;
      const float wz = wn * z;
      for (int j = 0; j < Ns; ++j) {
        dsum += __shfl_sync(0xffffffffu, wz, gbase + j);
        wsum += __shfl_sync(0xffffffffu, wn, gbase + j);
      }
      if (valid) {
        p.out_weights[(size_t)ray * Ns + k] = wn;
        if (k == 0) {
          if (p.white_bkgd) {
            const float bg = 1.f - wsum;
            ar += bg, ag += bg, ab += bg;
          }
          p.out_rgb[(size_t)ray * 3 + 0] = ar;
          p.out_rgb[(size_t)ray * 3 + 1] = ag;
          p.out_rgb[(size_t)ray * 3 + 2] = ab;
          p.out_depth[ray] = dsum;
        }
      }
    }
    // the next tile overwrites the A rows and the accumulators: all TMEM reads above are complete
    // (tcgen05.wait::ld) and G5 has been waited for; one CTA barrier orders the reuse
    RAY_STAMP(12);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    RAY_STAMP(13);
  }
  if (warp == 0) tc::tmem_dealloc(tmem, TMEM_COLS);
}

template <int S>
int launch_rays_tc(const RayTcParams& P, cudaStream_t stream) {
  constexpr size_t smem = TcSmem<S>::bytes;
  static PerDeviceSize attr_set;   // the attribute (and the SM count) is per device
  if (attr_set.cur() < smem) {
    cudaError_t e = cudaFuncSetAttribute(render_rays_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("render_rays_tc: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
      return ENERF_ECUDA;
    }
    attr_set.cur() = smem;
  }
  const int n_sm = device_sm_count();
  constexpr int per_sm = (S <= 4) ? 2 : 1;                          // 5-8 views: 512 TMEM columns / > 113 KB of shared memory per CTA
  const int grid = P.n_tiles < per_sm * n_sm ? P.n_tiles : per_sm * n_sm;  // persistent
  render_rays_tc_kernel<S><<<grid, 128, smem, stream>>>(P);
  ENERF_CHECK_LAUNCH("render_rays_tc");
  return ENERF_OK;
}

}  // namespace enerf
