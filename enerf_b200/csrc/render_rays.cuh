// render_rays.cuh -- the fused ray stage, FP32-pipe version (one thread = one ray).
//
// Replaces, in ONE launch and without any (N,Ns,S,C) intermediate in HBM
//   build_rays          /root/reference/lib/networks/enerf/utils.py:390-420
//   sample_along_depth  utils.py:422-441
//   get_vox_feat        utils.py:456-458   (trilinear, zeros padding, align_corners)
//   get_img_feat        utils.py:689-722   (bilinear, border padding, align_corners + ray-diff feats)
//   Agg.forward         lib/networks/enerf/nerf.py:74-89
//   NeRF.forward        nerf.py:29-43
//   raw2outputs         utils.py:571-603
//   (= Network.render_rays, lib/networks/enerf/network.py:24-43).
//
// The MLP weights (10,090 floats for fc=11; 14,050 for fc=35) are staged once per CTA into shared
// memory and read as warp-broadcast LDS.128.  The view-independent input columns of color.0
// (x, vox_img_feat: 88 of 88+fc+4) and of global_fc (var, mean: 2/3) are applied once per sample
// instead of once per view (SURVEY.md section 7 "MLP factorisation"); this only changes the fp32
// summation order.
#pragma once
#include "common.cuh"
#include "render_rays_params.cuh"

namespace enerf {


template <int FC>
struct RayW {                       // shared-memory layout (float offsets, all multiples of 4)
  static constexpr int VP = (FC + 4 + 3) / 4 * 4;   // padded per-view input width of color.0
  static constexpr int FCP = FC + 1;                // padded channel count of img_feat_rgb
  static constexpr int o_view_w = 0;                       // [4][FCP]
  static constexpr int o_view_b = o_view_w + 4 * FCP;      // [FCP]
  static constexpr int o_glob_w = o_view_b + FCP;          // [3*FC][32]
  static constexpr int o_glob_b = o_glob_w + 3 * FC * 32;  // [32]
  static constexpr int o_aggw_w = o_glob_b + 32;           // [32]
  static constexpr int o_aggw_b = o_aggw_w + 32;           // [4] (1 used)
  static constexpr int o_fc_w = o_aggw_b + 4;              // [32][16]
  static constexpr int o_fc_b = o_fc_w + 32 * 16;          // [16]
  static constexpr int o_lr0_w = o_fc_b + 16;              // [24][64]
  static constexpr int o_lr0_b = o_lr0_w + 24 * 64;        // [64]
  static constexpr int o_sig_w = o_lr0_b + 64;             // [64]
  static constexpr int o_sig_b = o_sig_w + 64;             // [4]
  static constexpr int o_c0s_w = o_sig_b + 4;              // [88][64]  shared columns (x, vif)
  static constexpr int o_c0_b = o_c0s_w + 88 * 64;         // [64]
  static constexpr int o_c0v_w = o_c0_b + 64;              // [64][VP]  per-view columns, transposed
  static constexpr int o_c2_w = o_c0v_w + 64 * VP;         // [64]
  static constexpr int o_c2_b = o_c2_w + 64;               // [4]
  static constexpr int total = o_c2_b + 4;
};

__device__ __forceinline__ void copy_to_smem(float* dst, const float* __restrict__ src, int n, int n_pad) {
  for (int e = threadIdx.x; e < n_pad; e += blockDim.x) dst[e] = (e < n) ? __ldg(src + e) : 0.f;
}

// out[0..N) += x * w[0..N)   (w in shared memory, 16-byte aligned, N % 4 == 0)
template <int N>
__device__ __forceinline__ void axpy_s(float* __restrict__ out, float x, const float* __restrict__ w) {
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(w + 4 * q);
    out[4 * q + 0] = fmaf(x, v.x, out[4 * q + 0]);
    out[4 * q + 1] = fmaf(x, v.y, out[4 * q + 1]);
    out[4 * q + 2] = fmaf(x, v.z, out[4 * q + 2]);
    out[4 * q + 3] = fmaf(x, v.w, out[4 * q + 3]);
  }
}

template <int FC, int SMAX, bool STATIC_S>
__global__ void __launch_bounds__(128) render_rays_kernel(const RayParams p) {
  using L = RayW<FC>;
  constexpr int FCP = L::FCP, VP = L::VP;
  extern __shared__ __align__(16) float sw[];
  __shared__ float cam_s[ENERF_MAX_VIEWS * 24 + 4];  // per view: ext(12) ixt(9) centre(3); then tar centre

  // ---- stage weights + cameras ----
  {
    // view_fc packed by the host as [4][FC] / [FC]; re-pitch to FCP so rows stay 16-byte aligned
    for (int e = threadIdx.x; e < 4 * FCP; e += blockDim.x) {
      const int j = e / FCP, c = e % FCP;
      sw[L::o_view_w + e] = (p.viewdir_agg && c < FC) ? __ldg(p.w[0] + j * FC + c) : 0.f;
    }
    for (int e = threadIdx.x; e < FCP; e += blockDim.x) sw[L::o_view_b + e] = (p.viewdir_agg && e < FC) ? __ldg(p.w[1] + e) : 0.f;
    copy_to_smem(sw + L::o_glob_w, p.w[2], 3 * FC * 32, 3 * FC * 32);
    copy_to_smem(sw + L::o_glob_b, p.w[3], 32, 32);
    copy_to_smem(sw + L::o_aggw_w, p.w[4], 32, 32);
    copy_to_smem(sw + L::o_aggw_b, p.w[5], 1, 4);
    copy_to_smem(sw + L::o_fc_w, p.w[6], 32 * 16, 32 * 16);
    copy_to_smem(sw + L::o_fc_b, p.w[7], 16, 16);
    copy_to_smem(sw + L::o_lr0_w, p.w[8], 24 * 64, 24 * 64);
    copy_to_smem(sw + L::o_lr0_b, p.w[9], 64, 64);
    copy_to_smem(sw + L::o_sig_w, p.w[10], 64, 64);
    copy_to_smem(sw + L::o_sig_b, p.w[11], 1, 4);
    copy_to_smem(sw + L::o_c0s_w, p.w[12], 88 * 64, 88 * 64);
    copy_to_smem(sw + L::o_c0_b, p.w[13], 64, 64);
    // color.0 per-view rows [88 .. 88+FC+4) of the [in][64] matrix -> transposed [64][VP]
    for (int e = threadIdx.x; e < 64 * VP; e += blockDim.x) {
      const int j = e / VP, i = e % VP;
      sw[L::o_c0v_w + e] = (i < FC + 4) ? __ldg(p.w[12] + (size_t)(88 + i) * 64 + j) : 0.f;
    }
    copy_to_smem(sw + L::o_c2_w, p.w[14], 64, 64);
    copy_to_smem(sw + L::o_c2_b, p.w[15], 1, 4);
    const EnerfCam* cam = p.cam;
    for (int e = threadIdx.x; e < p.S * 24; e += blockDim.x) {
      const int s = e / 24, k = e % 24;
      cam_s[e] = (k < 12) ? cam->src_ext[s][k] : (k < 21) ? cam->src_ixt[p.level][s][k - 12] : cam->src_center[s][k - 21];
    }
    if (threadIdx.x < 3) cam_s[ENERF_MAX_VIEWS * 24 + threadIdx.x] = cam->tar_center[threadIdx.x];
  }
  __syncthreads();

  int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= p.n_rays || (p.n_rays_dev != nullptr && ray >= __ldg(p.n_rays_dev))) return;
  if (p.out_raw) ray = (p.win_y + ray / p.win_w) * p.Wr + p.win_x + ray % p.win_w;   // window -> frame pixel
  const int S = STATIC_S ? SMAX : p.S;
  const int Ns = p.num_samples;

  // ---- build_rays: up-sampled depth / std / near_far at this ray's pixel, clamped interval ----
  const float4 r0 = ldg4(p.rays + (size_t)ray * 8), r1 = ldg4(p.rays + (size_t)ray * 8 + 4);
  const float ox = r0.x, oy = r0.y, oz = r0.z, dx = r0.w, dy = r1.x, dz = r1.y, u = r1.z, v = r1.w;
  const int ui = (int)u, vi = (int)v;  // .long() truncation, utils.py:416
  float dep, sd, vn, vf;
  {
    const size_t hw = (size_t)p.hv * p.wv;
    if (p.hv == p.Hr && p.wv == p.Wr) {
      const size_t o = (size_t)vi * p.wv + ui;
      dep = __ldg(p.depth + o), sd = __ldg(p.std + o), vn = __ldg(p.near_far + o), vf = __ldg(p.near_far + hw + o);
    } else {
      dep = bilinear_ac(p.depth, p.hv, p.wv, p.Hr, p.Wr, vi, ui);
      sd = bilinear_ac(p.std, p.hv, p.wv, p.Hr, p.Wr, vi, ui);
      vn = bilinear_ac(p.near_far, p.hv, p.wv, p.Hr, p.Wr, vi, ui);
      vf = bilinear_ac(p.near_far + hw, p.hv, p.wv, p.Hr, p.Wr, vi, ui);
    }
  }
  float rn, rf;
  if (p.depth_inv) {
    rn = fminf(dep + sd, vn);
    rf = fmaxf(dep - sd, vf);
  } else {
    rn = fmaxf(dep - sd, vn);
    rf = fminf(dep + sd, vf);
  }

  const float tcx = cam_s[ENERF_MAX_VIEWS * 24 + 0], tcy = cam_s[ENERF_MAX_VIEWS * 24 + 1], tcz = cam_s[ENERF_MAX_VIEWS * 24 + 2];
  const float un = u / (float)(p.Wr - 1), vnrm = v / (float)(p.Hr - 1);  // network.py:37

  float z_all[8], wgt_all[8];
  float T = 1.f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f;

#pragma unroll 1
  for (int k = 0; k < Ns; ++k) {
    // ---- sample_along_depth ----
    const float z = (Ns == 1) ? rn + (rf - rn) * 0.5f : rn + (rf - rn) * linspace01(k, Ns);
    const float tz = p.depth_inv ? 1.0f / fmaxf(z, 1e-6f) : z;
    const float X = ox + dx * tz, Y = oy + dy * tz, Z = oz + dz * tz;
    const float dn = p.depth_inv ? (vn - z) / fmaxf(vn - vf, 1e-6f) : (z - vn) / fmaxf(vf - vn, 1e-6f);

    // ---- get_vox_feat: trilinear, zeros padding ----
    float vif[24];
    {
      const float gx = un * 2.f - 1.f, gy = vnrm * 2.f - 1.f, gz = dn * 2.f - 1.f;
      const float ix = ((gx + 1.f) / 2.f) * (float)(p.wv - 1), iy = ((gy + 1.f) / 2.f) * (float)(p.hv - 1),
                  iz = ((gz + 1.f) / 2.f) * (float)(p.D - 1);
#pragma unroll
      for (int c = 0; c < 8; ++c) vif[c] = 0.f;
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
                vif[0] = fmaf(a.x, wgt, vif[0]), vif[1] = fmaf(a.y, wgt, vif[1]), vif[2] = fmaf(a.z, wgt, vif[2]),
                vif[3] = fmaf(a.w, wgt, vif[3]);
                vif[4] = fmaf(b.x, wgt, vif[4]), vif[5] = fmaf(b.y, wgt, vif[5]), vif[6] = fmaf(b.z, wgt, vif[6]),
                vif[7] = fmaf(b.w, wgt, vif[7]);
              }
            }
      }
    }

    // ---- get_img_feat: per-view projection, border-padded bilinear gather, ray-diff features ----
    float f[SMAX][FCP];   // [.., FC) image feature + rgb, slot FC unused padding
    float dir[SMAX][4];
    float tdx = X - tcx, tdy = Y - tcy, tdz = Z - tcz;
    {
      const float n = sqrtf(tdx * tdx + tdy * tdy + tdz * tdz) + 1e-6f;
      tdx /= n, tdy /= n, tdz /= n;
    }
#pragma unroll
    for (int s = 0; s < SMAX; ++s) {
      if (!STATIC_S && s >= S) break;
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
      const float gx = (p0 / pz) / (float)(p.Wr - 1) * 2.f - 1.f, gy = (p1 / pz) / (float)(p.Hr - 1) * 2.f - 1.f;
      float ix = ((gx + 1.f) / 2.f) * (float)(p.Wr - 1), iy = ((gy + 1.f) / 2.f) * (float)(p.Hr - 1);
      ix = fminf(fmaxf(ix, 0.f), (float)(p.Wr - 1));  // padding_mode='border': clip, then interpolate
      iy = fminf(fmaxf(iy, 0.f), (float)(p.Hr - 1));
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      const int x0 = (int)fx0, y0 = (int)fy0;
      const int x1 = min(x0 + 1, p.Wr - 1), y1 = min(y0 + 1, p.Hr - 1);  // weight is 0 whenever clamped
      const float txr = (fx0 + 1.f) - ix, txl = ix - fx0, tyb = (fy0 + 1.f) - iy, tyt = iy - fy0;
      const float w_nw = txr * tyb, w_ne = txl * tyb, w_sw = txr * tyt, w_se = txl * tyt;
      const float* base = p.img + (size_t)s * p.Hr * p.Wr * FCP;
      const float* q00 = base + ((size_t)y0 * p.Wr + x0) * FCP;
      const float* q01 = base + ((size_t)y0 * p.Wr + x1) * FCP;
      const float* q10 = base + ((size_t)y1 * p.Wr + x0) * FCP;
      const float* q11 = base + ((size_t)y1 * p.Wr + x1) * FCP;
#pragma unroll
      for (int q = 0; q < FCP / 4; ++q) {
        const float4 a = ldg4(q00 + 4 * q), b = ldg4(q01 + 4 * q), c = ldg4(q10 + 4 * q), d = ldg4(q11 + 4 * q);
        f[s][4 * q + 0] = a.x * w_nw + b.x * w_ne + c.x * w_sw + d.x * w_se;
        f[s][4 * q + 1] = a.y * w_nw + b.y * w_ne + c.y * w_sw + d.y * w_se;
        f[s][4 * q + 2] = a.z * w_nw + b.z * w_ne + c.z * w_sw + d.z * w_se;
        f[s][4 * q + 3] = a.w * w_nw + b.w * w_ne + c.w * w_sw + d.w * w_se;
      }
      float sx = X - Cn[0], sy = Y - Cn[1], sz = Z - Cn[2];
      const float n = sqrtf(sx * sx + sy * sy + sz * sz) + 1e-6f;
      sx /= n, sy /= n, sz /= n;
      const float rx = tdx - sx, ry = tdy - sy, rz = tdz - sz;
      const float rnm = fmaxf(sqrtf(rx * rx + ry * ry + rz * rz), 1e-6f);
      dir[s][0] = rx / rnm, dir[s][1] = ry / rnm, dir[s][2] = rz / rnm;
      dir[s][3] = tdx * sx + tdy * sy + tdz * sz;
    }

    // ---- Agg: view_fc, var/mean over views, global_fc, softmax-weighted pooling, fc ----
    {
      float g[SMAX][FCP];
#pragma unroll
      for (int s = 0; s < SMAX; ++s) {
        if (!STATIC_S && s >= S) break;
        float t[FCP];
#pragma unroll
        for (int c = 0; c < FCP; ++c) t[c] = sw[L::o_view_b + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) axpy_s<FCP>(t, dir[s][j], sw + L::o_view_w + j * FCP);
#pragma unroll
        for (int c = 0; c < FCP; ++c) g[s][c] = f[s][c] + (p.viewdir_agg ? fmaxf(t[c], 0.f) : 0.f);
      }
      float hb[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) hb[j] = sw[L::o_glob_b + j];
      const float invS = 1.0f / (float)S, invS1 = 1.0f / (float)(S - 1);
#pragma unroll
      for (int c = 0; c < FC; ++c) {
        float m = 0.f;
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
          if (STATIC_S || s < S) m += g[s][c];
        m *= invS;
        float vr = 0.f;
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
          if (STATIC_S || s < S) vr += (g[s][c] - m) * (g[s][c] - m);
        vr *= invS1;  // unbiased (torch.var default, nerf.py:82)
        axpy_s<32>(hb, vr, sw + L::o_glob_w + (FC + c) * 32);
        axpy_s<32>(hb, m, sw + L::o_glob_w + (2 * FC + c) * 32);
      }
      // per-view hidden + attention logits; softmax over views exactly as torch (max-subtracted)
      float im[32];
      float lg[SMAX];
      if constexpr (STATIC_S && SMAX <= 4 && FC <= 12) {
        float h[SMAX][32];
#pragma unroll
        for (int s = 0; s < SMAX; ++s)
#pragma unroll
          for (int j = 0; j < 32; ++j) h[s][j] = hb[j];
#pragma unroll
        for (int c = 0; c < FC; ++c) {
          float wr[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(sw + L::o_glob_w + c * 32 + 4 * q);
            wr[4 * q] = v4.x, wr[4 * q + 1] = v4.y, wr[4 * q + 2] = v4.z, wr[4 * q + 3] = v4.w;
          }
#pragma unroll
          for (int s = 0; s < SMAX; ++s)
#pragma unroll
            for (int j = 0; j < 32; ++j) h[s][j] = fmaf(g[s][c], wr[j], h[s][j]);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int s = 0; s < SMAX; ++s) {
          float a = sw[L::o_aggw_b];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            h[s][j] = fmaxf(h[s][j], 0.f);
            a = fmaf(h[s][j], sw[L::o_aggw_w + j], a);
          }
          lg[s] = fmaxf(a, 0.f);
          mx = fmaxf(mx, lg[s]);
        }
        float den = 0.f;
#pragma unroll
        for (int s = 0; s < SMAX; ++s) {
          lg[s] = expf(lg[s] - mx);
          den += lg[s];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) im[j] = 0.f;
#pragma unroll
        for (int s = 0; s < SMAX; ++s) {
          const float ws_ = lg[s] / den;
#pragma unroll
          for (int j = 0; j < 32; ++j) im[j] = fmaf(h[s][j], ws_, im[j]);
        }
      } else {
        // generic path: two passes over the views (logits, then weighted sum), recomputing h
        float mx = -INFINITY;
        for (int s = 0; s < S; ++s) {
          float h[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = hb[j];
#pragma unroll 1
          for (int c = 0; c < FC; ++c) axpy_s<32>(h, g[s][c], sw + L::o_glob_w + c * 32);
          float a = sw[L::o_aggw_b];
#pragma unroll
          for (int j = 0; j < 32; ++j) a = fmaf(fmaxf(h[j], 0.f), sw[L::o_aggw_w + j], a);
          lg[s] = fmaxf(a, 0.f);
          mx = fmaxf(mx, lg[s]);
        }
        float den = 0.f;
        for (int s = 0; s < S; ++s) {
          lg[s] = expf(lg[s] - mx);
          den += lg[s];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) im[j] = 0.f;
        for (int s = 0; s < S; ++s) {
          float h[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = hb[j];
#pragma unroll 1
          for (int c = 0; c < FC; ++c) axpy_s<32>(h, g[s][c], sw + L::o_glob_w + c * 32);
          const float ws_ = lg[s] / den;
#pragma unroll
          for (int j = 0; j < 32; ++j) im[j] = fmaf(fmaxf(h[j], 0.f), ws_, im[j]);
        }
      }
      // fc 32 -> 16, ReLU  => vox_img_feat[8..24)
#pragma unroll
      for (int c = 0; c < 16; ++c) vif[8 + c] = sw[L::o_fc_b + c];
#pragma unroll
      for (int j = 0; j < 32; ++j) axpy_s<16>(vif + 8, im[j], sw + L::o_fc_w + j * 16);
#pragma unroll
      for (int c = 0; c < 16; ++c) vif[8 + c] = fmaxf(vif[8 + c], 0.f);
    }

    // ---- NeRF: lr0, sigma, color ----
    float sigma, cb[64];
    {
      float x[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) x[j] = sw[L::o_lr0_b + j];
#pragma unroll
      for (int i = 0; i < 24; ++i) axpy_s<64>(x, vif[i], sw + L::o_lr0_w + i * 64);
      float sg = sw[L::o_sig_b];
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        x[j] = fmaxf(x[j], 0.f);
        sg = fmaf(x[j], sw[L::o_sig_w + j], sg);
      }
      sigma = (sg > 20.f) ? sg : log1pf(expf(sg));  // nn.Softplus(beta=1, threshold=20)
#pragma unroll
      for (int j = 0; j < 64; ++j) cb[j] = sw[L::o_c0_b + j];
#pragma unroll
      for (int i = 0; i < 64; ++i) axpy_s<64>(cb, x[i], sw + L::o_c0s_w + i * 64);
#pragma unroll
      for (int i = 0; i < 24; ++i) axpy_s<64>(cb, vif[i], sw + L::o_c0s_w + (64 + i) * 64);
    }
    float cl[SMAX];
#pragma unroll
    for (int s = 0; s < SMAX; ++s) cl[s] = sw[L::o_c2_b];
#pragma unroll 4
    for (int j = 0; j < 64; ++j) {
      const float* wr = sw + L::o_c0v_w + j * VP;
      const float w2 = sw[L::o_c2_w + j];
#pragma unroll
      for (int s = 0; s < SMAX; ++s) {
        if (!STATIC_S && s >= S) break;
        float t = cb[j];
#pragma unroll
        for (int q = 0; q < VP / 4; ++q) {
          const float4 v4 = *reinterpret_cast<const float4*>(wr + 4 * q);
          const float wv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * q + e;
            if (i < FC) t = fmaf(f[s][i], wv[e], t);
            else if (i < FC + 4) t = fmaf(dir[s][i - FC], wv[e], t);
          }
        }
        cl[s] = fmaf(fmaxf(t, 0.f), w2, cl[s]);
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < SMAX; ++s)
      if (STATIC_S || s < S) {
        cl[s] = fmaxf(cl[s], 0.f);
        mx = fmaxf(mx, cl[s]);
      }
    float den = 0.f;
#pragma unroll
    for (int s = 0; s < SMAX; ++s)
      if (STATIC_S || s < S) {
        cl[s] = expf(cl[s] - mx);
        den += cl[s];
      }
    float cr = 0.f, cg = 0.f, cbl = 0.f;
#pragma unroll
    for (int s = 0; s < SMAX; ++s)
      if (STATIC_S || s < S) {
        const float ws_ = cl[s] / den;
        cr = fmaf(f[s][FC - 3], ws_, cr), cg = fmaf(f[s][FC - 2], ws_, cg), cbl = fmaf(f[s][FC - 1], ws_, cbl);
      }

    if (p.out_raw) {   // layered mode: the samples are merged across layers later (enerf_composite_layers)
      const size_t o = (size_t)ray * p.out_stride + p.out_off + k;
      *reinterpret_cast<float4*>(p.out_raw + o * 4) = make_float4(cr, cg, cbl, sigma);
      p.out_z[o] = p.depth_inv ? 1.0f / z : z;   // network_composite.py:48-51
      continue;
    }
    // ---- raw2outputs (running transmittance) ----
    const float alpha = 1.f - expf(-sigma);
    const float wk = alpha * T;
    T *= (1.f - alpha + 1e-10f);
    acc_r = fmaf(wk, cr, acc_r), acc_g = fmaf(wk, cg, acc_g), acc_b = fmaf(wk, cbl, acc_b);
    z_all[k] = z;
    wgt_all[k] = wk;
  }

  if (p.out_raw) return;
  // depth uses softmax(weights) (utils.py:594-595)
  float mx = -INFINITY;
  for (int k = 0; k < Ns; ++k) mx = fmaxf(mx, wgt_all[k]);
  float den = 0.f;
  for (int k = 0; k < Ns; ++k) {
    wgt_all[k] = expf(wgt_all[k] - mx);
    den += wgt_all[k];
  }
  float dsum = 0.f, wsum = 0.f;
  for (int k = 0; k < Ns; ++k) {
    const float wn = wgt_all[k] / den;
    p.out_weights[(size_t)ray * Ns + k] = wn;
    dsum += wn * z_all[k];
    wsum += wn;
  }
  if (p.white_bkgd) {
    const float bg = 1.f - wsum;
    acc_r += bg, acc_g += bg, acc_b += bg;
  }
  p.out_rgb[(size_t)ray * 3 + 0] = acc_r;
  p.out_rgb[(size_t)ray * 3 + 1] = acc_g;
  p.out_rgb[(size_t)ray * 3 + 2] = acc_b;
  p.out_depth[ray] = dsum;
}

template <int FC, int SMAX, bool STATIC_S>
int launch_rays(const RayParams& p, cudaStream_t stream) {
  constexpr size_t smem = RayW<FC>::total * sizeof(float);
  static PerDeviceSize attr_set;   // the attribute is per device
  if (attr_set.cur() < smem) {
    cudaError_t e = cudaFuncSetAttribute(render_rays_kernel<FC, SMAX, STATIC_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("render_rays: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
      return ENERF_ECUDA;
    }
    attr_set.cur() = smem;
  }
  render_rays_kernel<FC, SMAX, STATIC_S><<<ceil_div(p.n_rays, 128), 128, smem, stream>>>(p);
  ENERF_CHECK_LAUNCH("render_rays");
  return ENERF_OK;
}

}  // namespace enerf
