// render_rays_ws.cu -- the fused ray stage, warp-specialised (the default tensor-core path for 2 or 3 source views).
//
// Same scope and arithmetic as render_rays_tc.cu (build_rays + sample_along_depth + get_vox_feat + get_img_feat +
// Agg + NeRF + raw2outputs; /root/reference/lib/networks/enerf/network.py:24-43, nerf.py:29-43,74-89,
// utils.py:390-441,456-458,571-603,689-722), re-organised around what the round-1 profile showed: the kernel was
// latency-bound in its gather phase (52 scattered 16-byte loads per sample point at 8 warps / SM) and the tensor
// pipe idled meanwhile.  Here ONE persistent CTA per SM runs three warpgroups:
//
//   warps 4-7, 8-11  two GATHER groups.  Group g owns every second tile of the CTA: thread j gathers sample point j
//                    (trilinear voxel feature, S bilinear image features, ray-diff features, view_fc / var / mean on
//                    the FP32 pipe) and writes its row of the A operands into the group's own buffer.
//   warps 0-3        the CONSUMER: issues the tcgen05.mma batches (warp 0, one elected lane) and runs the epilogues
//                    out of TMEM (128 threads = 128 TMEM lanes = 128 sample points), then composites the samples
//                    of a ray with warp shuffles.
//
// so the gathers of tiles i+1 and i+2 run under the GEMMs / epilogues of tile i.  Every GEMM whose A operand is
// gathered data (global_fc, the per-view and voxel columns of color.0 and lr0) is issued in ONE batch at the start
// of a tile and releases the gather buffer as soon as it has been read.
//
// MLP factorisation: the view-independent input columns are contracted ONCE per point --
// color.0 = W[:, :88] [x | vox | img]  (shared, 11 K-steps)  +  W[:, 88:] [f_s | dir_s]  (per view, 2 K-steps):
// 17 MMAs instead of 39; global_fc = W[:, fc:] [var | mean] (3) + W[:, :fc] g_s (2 per view): 9 instead of 15.
// The epilogues add the two accumulators.  33 MMAs per 128-point tile instead of 61.
//
// Numerics: exactly render_rays_tc.cu's (TF32 operands via cvt.rna, fp32 accumulate, fp32 everywhere else); the
// factorisation only changes the order of the fp32 accumulation inside the tensor core.
#include "common.cuh"
#include "render_rays_params.cuh"
#include "tc.cuh"

namespace enerf {

namespace ws {

constexpr int FC = 11;         // image feature (8) + rgb (3)
constexpr int FCP = 12;

// float offsets inside the packed weight blob (host: enerf_b200/packing.py::pack_nerf_tc) -- same blob as render_rays_tc.cu
struct W {
  static constexpr int bg_view = 0;                      // B [4 chunks][32][4]   K=16 N=32  (rows 0-10 g, 11 bias)
  static constexpr int bg_shared = bg_view + 4 * 32 * 4; // B [6][32][4]          K=24 N=32  ([var | mean])
  static constexpr int bfc = bg_shared + 6 * 32 * 4;     // B [8][16][4]          K=32 N=16
  static constexpr int b0 = bfc + 8 * 16 * 4;            // B [6][64][4]          K=24 N=64  ([vox 8 | img 16])
  static constexpr int bc_shared = b0 + 6 * 64 * 4;      // B [22][64][4]         K=88 N=64  ([x 64 | vox 8 | img 16])
  static constexpr int bc_view = bc_shared + 22 * 64 * 4;  // B [4][64][4]        K=16 N=64  ([f 11 | dir 4 | bias])
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

constexpr int CHUNK = 512;     // floats: 128 rows x 16 bytes (K-major, no swizzle: row r of chunk c at c*2048 + r*16)

template <int S>
struct Smem {
  // gather buffer (one per gather group), in chunks
  static constexpr int c_fd = 0;                 // [f | dir | 1]: 4 per view  (A of color.0's per-view columns)
  static constexpr int c_g = 4 * S;              // g_s = f_s + relu(view_fc(dir_s)), [.. | 1]: 3 per view (A of global_fc's per-view columns)
  static constexpr int c_zero = c_g + 3 * S;     // one chunk of zeros: the 4th chunk of every g_s (K = 16) through the descriptor's LBO
  static constexpr int c_vm = c_zero + 1;        // [var (12) | mean (12)]: 6
  static constexpr int c_vox = c_vm + 6;         // voxel feature: 2
  static constexpr int G_CHUNKS = c_vox + 2;
  static constexpr int N_SCAL = 3 * S + 1;       // per point, fp32: the S source colours (the blend must not see TF32 rounding), z
  static constexpr int g_floats = G_CHUNKS * CHUNK + N_SCAL * 128;
  // consumer region P: x (16 chunks) | img (4); im (8 chunks) aliases the first half of x
  static constexpr int p_x = 0, p_img = 16, P_CHUNKS = 20;
  static constexpr int o_w = 0;
  static constexpr int o_g = (W::total + 31) / 32 * 32;
  static constexpr int o_p = o_g + 2 * g_floats;
  static constexpr int o_cam = o_p + P_CHUNKS * CHUNK;
  static constexpr int total_floats = o_cam + ENERF_MAX_VIEWS * 24 + 4;
  static constexpr size_t bytes = (size_t)total_floats * 4 + 128;
  // TMEM columns
  static constexpr int t_g1s = 0;                // global_fc, shared part (32); fc's 16 columns reuse it after E1
  static constexpr int t_g1v = 32;               // global_fc, per view (32 each)
  static constexpr int t_lr0 = 128;              // lr0 (64)
  static constexpr int t_cs = 192;               // color.0 shared part (64)
  static constexpr int t_cv = 256;               // color.0 per view (64 each)
  static constexpr int tmem_cols = 512;
  static_assert(32 + 32 * S <= 128 && 256 + 64 * S <= 512, "TMEM layout holds S <= 3... 4 views");
};

__device__ __forceinline__ void store_chunk(float* base, int chunk, int row, float v0, float v1, float v2, float v3) {
  *reinterpret_cast<float4*>(base + (size_t)chunk * CHUNK + row * 4) = make_float4(tc::to_tf32(v0), tc::to_tf32(v1), tc::to_tf32(v2), tc::to_tf32(v3));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct Params {
  RayParams r;
  const float* wblob;
  int n_tiles;
  unsigned long long* dbg;    // optional %globaltimer stamps of CTA 0 (enerf_render_rays_debug): [role 0..2][tile 0..7][16]
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// role 0 = consumer (thread 0), 1 / 2 = gather group 0 / 1 (their row 0); tile = the role's own tile counter
#define WS_STAMP(role_, tile_, i_)                                                                   \
  do {                                                                                               \
    if (dbg && (tile_) < 8) dbg[((role_) * 8 + (tile_)) * 16 + (i_)] = gtime();                      \
  } while (0)

template <int S>
__global__ void __launch_bounds__(384, 1) render_rays_ws_kernel(const Params P) {
  using SM = Smem<S>;
  extern __shared__ __align__(128) float smem[];
  float* sw = smem + SM::o_w;
  float* p_s = smem + SM::o_p;
  float* cam_s = smem + SM::o_cam;
  __shared__ __align__(8) uint64_t full_bar[2], free_bar[2], mma_bar;
  __shared__ uint32_t tmem_base_s;
  const RayParams& p = P.r;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  // ---- one-time setup: weights, cameras, zero chunks, barriers, TMEM ----
  for (int e = t; e < W::total; e += 384) sw[e] = __ldg(P.wblob + e);
  {
    const EnerfCam* cam = p.cam;
    for (int e = t; e < S * 24; e += 384) {
      const int s = e / 24, k = e % 24;
      cam_s[e] = (k < 12) ? cam->src_ext[s][k] : (k < 21) ? cam->src_ixt[p.level][s][k - 12] : cam->src_center[s][k - 21];
    }
    if (t < 3) cam_s[ENERF_MAX_VIEWS * 24 + t] = cam->tar_center[t];
    for (int e = t; e < 2 * CHUNK; e += 384) smem[SM::o_g + (e / CHUNK) * SM::g_floats + SM::c_zero * CHUNK + (e % CHUNK)] = 0.f;
  }
  if (t == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&full_bar[i], 128);     // every thread of the gather group
      tc::mbar_init(&free_bar[i], 1);       // tcgen05.commit of the batch that read the buffer
    }
    tc::mbar_init(&mma_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, SM::tmem_cols);
  tc::fence_proxy_async();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  const int Ns = p.num_samples;
  // device-side ray count (masked rays without a host read-back): the grid was sized for the upper bound p.n_rays
  const int n_rays = p.n_rays_dev ? min(__ldg(p.n_rays_dev), p.n_rays) : p.n_rays;
  const int n_tiles = p.n_rays_dev ? (int)(((long long)n_rays * Ns + 127) / 128) : P.n_tiles;
  const float tcx = cam_s[ENERF_MAX_VIEWS * 24 + 0], tcy = cam_s[ENERF_MAX_VIEWS * 24 + 1], tcz = cam_s[ENERF_MAX_VIEWS * 24 + 2];
  const size_t hw = (size_t)p.hv * p.wv;

  if (warp >= 4) {
    // ======================================= gather groups =======================================
    const int grp = (warp - 4) >> 2;                 // 0 | 1
    const int row = t - 128 - 128 * grp;             // 0..127: the sample point of the tile this thread gathers
    float* gb = smem + SM::o_g + grp * SM::g_floats;
    float* scal = gb + SM::G_CHUNKS * CHUNK;
    int j = 0;                                       // this group's tile counter
    unsigned long long* dbg = (P.dbg && blockIdx.x == 0 && row == 0) ? P.dbg : nullptr;
    for (int k = grp, tile = blockIdx.x + grp * gridDim.x; tile < n_tiles; k += 2, tile += 2 * gridDim.x, ++j) {
      WS_STAMP(1 + grp, j, 0);
      const long long pt = (long long)tile * 128 + row;
      const bool valid = pt < (long long)n_rays * Ns;
      int ray = valid ? (int)(pt / Ns) : n_rays - 1;
      if (p.out_raw) ray = (p.win_y + ray / p.win_w) * p.Wr + p.win_x + ray % p.win_w;   // layered mode: window -> frame pixel
      const int ks = (int)(pt % Ns);

      // ---- build_rays + sample_along_depth ----
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
      const float z = (Ns == 1) ? rn + (rf - rn) * 0.5f : rn + (rf - rn) * linspace01(ks, Ns);
      const float tz = p.depth_inv ? 1.0f / fmaxf(z, 1e-6f) : z;
      const float X = r0.x + r0.w * tz, Y = r0.y + r1.x * tz, Z = r0.z + r1.y * tz;
      const float dn = p.depth_inv ? (vn - z) / fmaxf(vn - vf, 1e-6f) : (z - vn) / fmaxf(vf - vn, 1e-6f);

      // ---- get_vox_feat: trilinear, zeros padding ----
      float vox[8];
      {
        const float un = u / (float)(p.Wr - 1), vnrm = v / (float)(p.Hr - 1);
        const float gx = un * 2.f - 1.f, gy = vnrm * 2.f - 1.f, gz = dn * 2.f - 1.f;
        const float ix = ((gx + 1.f) / 2.f) * (float)(p.wv - 1), iy = ((gy + 1.f) / 2.f) * (float)(p.hv - 1),
                    iz = ((gz + 1.f) / 2.f) * (float)(p.D - 1);
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

      // ---- get_img_feat: per-view projection, border-padded bilinear gather, ray-diff features ----
      float f[S][FCP], dir[S][4];
      float tdx = X - tcx, tdy = Y - tcy, tdz = Z - tcz;
      {
        const float n = sqrtf(tdx * tdx + tdy * tdy + tdz * tdz) + 1e-6f;
        tdx /= n, tdy /= n, tdz /= n;
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
        const float gx = (p0 / pz) / (float)(p.Wr - 1) * 2.f - 1.f, gy = (p1 / pz) / (float)(p.Hr - 1) * 2.f - 1.f;
        float ix = ((gx + 1.f) / 2.f) * (float)(p.Wr - 1), iy = ((gy + 1.f) / 2.f) * (float)(p.Hr - 1);
        ix = fminf(fmaxf(ix, 0.f), (float)(p.Wr - 1));
        iy = fminf(fmaxf(iy, 0.f), (float)(p.Hr - 1));
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        const int x1 = min(x0 + 1, p.Wr - 1), y1 = min(y0 + 1, p.Hr - 1);
        const float txr = (fx0 + 1.f) - ix, txl = ix - fx0, tyb = (fy0 + 1.f) - iy, tyt = iy - fy0;
        const float w_nw = txr * tyb, w_ne = txl * tyb, w_sw = txr * tyt, w_se = txl * tyt;
        const float* base = p.img + (size_t)s * p.Hr * p.Wr * FCP;
        const float* q00 = base + ((size_t)y0 * p.Wr + x0) * FCP;
        const float* q01 = base + ((size_t)y0 * p.Wr + x1) * FCP;
        const float* q10 = base + ((size_t)y1 * p.Wr + x0) * FCP;
        const float* q11 = base + ((size_t)y1 * p.Wr + x1) * FCP;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
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
        dir[s][0] = rx / rnm, dir[s][1] = ry / rnm, dir[s][2] = rz / rnm, dir[s][3] = tdx * sx + tdy * sy + tdz * sz;
      }

      // ---- the buffer is free once the consumer's first MMA batch of this group's previous tile has read it ----
      WS_STAMP(1 + grp, j, 1);     // gathers done (in registers)
      tc::mbar_wait(&free_bar[grp], (uint32_t)((j & 1) ^ 1));
      WS_STAMP(1 + grp, j, 2);     // buffer free
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int cf = SM::c_fd + 4 * s;
        store_chunk(gb, cf + 0, row, f[s][0], f[s][1], f[s][2], f[s][3]);
        store_chunk(gb, cf + 1, row, f[s][4], f[s][5], f[s][6], f[s][7]);
        store_chunk(gb, cf + 2, row, f[s][8], f[s][9], f[s][10], dir[s][0]);
        store_chunk(gb, cf + 3, row, dir[s][1], dir[s][2], dir[s][3], 1.f);   // column 15 = 1: color.0's bias rides in B (row 15)
        scal[(3 * s + 0) * 128 + row] = f[s][8], scal[(3 * s + 1) * 128 + row] = f[s][9], scal[(3 * s + 2) * 128 + row] = f[s][10];
        // g_s = f_s + relu(view_fc(dir_s))    (nerf.py:76-78)
        if (p.viewdir_agg) {
#pragma unroll
          for (int c = 0; c < FC; ++c) {
            float a = sw[W::v_view_b + c];
            a = fmaf(dir[s][0], sw[W::v_view_w + 0 * 12 + c], a);
            a = fmaf(dir[s][1], sw[W::v_view_w + 1 * 12 + c], a);
            a = fmaf(dir[s][2], sw[W::v_view_w + 2 * 12 + c], a);
            a = fmaf(dir[s][3], sw[W::v_view_w + 3 * 12 + c], a);
            f[s][c] += fmaxf(a, 0.f);
          }
        }
        const int cg = SM::c_g + 3 * s;
        store_chunk(gb, cg + 0, row, f[s][0], f[s][1], f[s][2], f[s][3]);
        store_chunk(gb, cg + 1, row, f[s][4], f[s][5], f[s][6], f[s][7]);
        store_chunk(gb, cg + 2, row, f[s][8], f[s][9], f[s][10], 1.f);        // column 11 = 1: global_fc's bias rides in B
      }
      {
        float vr[12], mn[12];
        vr[11] = 0.f, mn[11] = 0.f;
        const float invS = 1.0f / (float)S, invS1 = 1.0f / (float)(S - 1);
#pragma unroll
        for (int c = 0; c < FC; ++c) {
          float m = 0.f;
#pragma unroll
          for (int s = 0; s < S; ++s) m += f[s][c];
          m *= invS;
          float q = 0.f;
#pragma unroll
          for (int s = 0; s < S; ++s) q += (f[s][c] - m) * (f[s][c] - m);
          vr[c] = q * invS1, mn[c] = m;
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          store_chunk(gb, SM::c_vm + q, row, vr[4 * q], vr[4 * q + 1], vr[4 * q + 2], vr[4 * q + 3]);
          store_chunk(gb, SM::c_vm + 3 + q, row, mn[4 * q], mn[4 * q + 1], mn[4 * q + 2], mn[4 * q + 3]);
        }
      }
      store_chunk(gb, SM::c_vox + 0, row, vox[0], vox[1], vox[2], vox[3]);
      store_chunk(gb, SM::c_vox + 1, row, vox[4], vox[5], vox[6], vox[7]);
      scal[(3 * S) * 128 + row] = z;
      tc::fence_proxy_async();             // my rows -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[grp]);
      WS_STAMP(1 + grp, j, 3);     // rows stored
    }
  } else {
    // ======================================= consumer =======================================
    const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);       // warp-uniform copy for the MMA issue
    const uint32_t w_addr = tc::smem_u32(sw), p_addr = tc::smem_u32(p_s);
    uint32_t phase = 0;
    auto desc_a = [&](uint32_t base, int chunk, uint32_t lbo) { return tc::smem_desc(base + (uint32_t)chunk * 2048u, lbo, 128u); };
    auto desc_b = [&](int off_floats, int chunk, int N) {
      return tc::smem_desc(w_addr + (uint32_t)off_floats * 4u + (uint32_t)chunk * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
    };
    auto publish = [&]() {   // A rows written by the 128 consumer threads -> visible to the tensor core, then everyone is past
      tc::fence_proxy_async();
      tc::tc_fence_before_sync();
      consumer_sync();
      tc::tc_fence_after_sync();
    };
    auto wait_mma = [&]() {
      tc::mbar_wait(&mma_bar, phase);
      phase ^= 1;
      tc::tc_fence_after_sync();
    };
    int k = 0;
    unsigned long long* dbg = (P.dbg && blockIdx.x == 0 && t == 0) ? P.dbg : nullptr;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
      const int b = k & 1;
      WS_STAMP(0, k, 0);
      const uint32_t g_addr = tc::smem_u32(smem + SM::o_g + b * SM::g_floats);
      const float* scal = smem + SM::o_g + b * SM::g_floats + SM::G_CHUNKS * CHUNK;
      const long long pt = (long long)tile * 128 + t;
      const bool valid = pt < (long long)n_rays * Ns;
      int ray = valid ? (int)(pt / Ns) : n_rays - 1;
      if (p.out_raw) ray = (p.win_y + ray / p.win_w) * p.Wr + p.win_x + ray % p.win_w;
      const int ks = (int)(pt % Ns);

      tc::mbar_wait(&full_bar[b], (uint32_t)((k >> 1) & 1));
      WS_STAMP(0, k, 1);           // gathered tile available
      float rgb_s[S][3];
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb_s[s][c] = scal[(3 * s + c) * 128 + t];
      const float z = scal[(3 * S) * 128 + t];
      // everyone has read its scalars (the buffer is released by the batch below), and the previous tile's TMEM reads are done
      tc::tc_fence_before_sync();
      consumer_sync();
      tc::tc_fence_after_sync();

      // ===== batch 1: every GEMM fed by gathered data =====
      if (warp == 0) {
        const uint32_t id32 = tc::idesc_tf32(128, 32), id64 = tc::idesc_tf32(128, 64);
        // global_fc, shared columns: [var | mean] (K = 24)
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
          tc::mma_tf32_elect(tmem_u + SM::t_g1s, desc_a(g_addr, SM::c_vm + 2 * kk, 2048u), desc_b(W::bg_shared, 2 * kk, 32), id32, kk > 0);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          // global_fc, per-view columns: g_s (K = 16; the 4th chunk is the shared zero chunk, reached through LBO)
          tc::mma_tf32_elect(tmem_u + SM::t_g1v + 32 * s, desc_a(g_addr, SM::c_g + 3 * s, 2048u), desc_b(W::bg_view, 0, 32), id32, 0);
          tc::mma_tf32_elect(tmem_u + SM::t_g1v + 32 * s, desc_a(g_addr, SM::c_g + 3 * s + 2, (uint32_t)(SM::c_zero - (SM::c_g + 3 * s + 2)) * 2048u),
                             desc_b(W::bg_view, 2, 32), id32, 1);
          // color.0, per-view columns: [f_s | dir_s | 1] (K = 16)
          tc::mma_tf32_elect(tmem_u + SM::t_cv + 64 * s, desc_a(g_addr, SM::c_fd + 4 * s, 2048u), desc_b(W::bc_view, 0, 64), id64, 0);
          tc::mma_tf32_elect(tmem_u + SM::t_cv + 64 * s, desc_a(g_addr, SM::c_fd + 4 * s + 2, 2048u), desc_b(W::bc_view, 2, 64), id64, 1);
        }
        // voxel columns of color.0's shared part (rows 64..71 of W) and of lr0 (rows 0..7)
        tc::mma_tf32_elect(tmem_u + SM::t_cs, desc_a(g_addr, SM::c_vox, 2048u), desc_b(W::bc_shared, 16, 64), id64, 0);
        tc::mma_tf32_elect(tmem_u + SM::t_lr0, desc_a(g_addr, SM::c_vox, 2048u), desc_b(W::b0, 0, 64), id64, 0);
        tc::mma_commit_elect(&free_bar[b]);      // the gather group may refill the buffer once these MMAs have read it
        tc::mma_commit_elect(&mma_bar);
        __syncwarp();
      }
      WS_STAMP(0, k, 2);           // batch 1 issued
      wait_mma();
      WS_STAMP(0, k, 3);           // batch 1 complete
      // ---- E1: global_fc = shared + view, ReLU, agg_w_fc logits, softmax over views, weighted pooling (nerf.py:85-88) ----
      {
        float sh[32], im[32];
        tc::tmem_ld32(tmem_row + SM::t_g1s, sh);
        tc::tmem_ld_wait();
#pragma unroll
        for (int jx = 0; jx < 32; ++jx) im[jx] = 0.f;
        float mx = -INFINITY, den = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float h[32];
          tc::tmem_ld32(tmem_row + SM::t_g1v + 32 * s, h);
          tc::tmem_ld_wait();
          float a = sw[W::v_ba];
#pragma unroll
          for (int jx = 0; jx < 32; ++jx) {
            h[jx] = fmaxf(sh[jx] + h[jx], 0.f);                   // bias already in the per-view accumulator
            a = fmaf(h[jx], sw[W::v_wa + jx], a);
          }
          const float lg = fmaxf(a, 0.f);
          const float mnew = fmaxf(mx, lg);
          const float sc = expf(mx - mnew), e = expf(lg - mnew);   // first view: sc = exp(-inf) = 0
          den = fmaf(den, sc, e);
#pragma unroll
          for (int jx = 0; jx < 32; ++jx) im[jx] = fmaf(e, h[jx], im[jx] * sc);
          mx = mnew;
        }
        const float inv = 1.0f / den;
#pragma unroll
        for (int q = 0; q < 8; ++q) store_chunk(p_s, SM::p_x + q, t, im[4 * q] * inv, im[4 * q + 1] * inv, im[4 * q + 2] * inv, im[4 * q + 3] * inv);
      }
      WS_STAMP(0, k, 4);           // E1 done
      publish();
      // ===== G2: fc 32 -> 16 =====
      if (warp == 0) {
        const uint32_t id = tc::idesc_tf32(128, 16);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc::mma_tf32_elect(tmem_u + SM::t_g1s, desc_a(p_addr, SM::p_x + 2 * kk, 2048u), desc_b(W::bfc, 2 * kk, 16), id, kk > 0);
        tc::mma_commit_elect(&mma_bar);
        __syncwarp();
      }
      wait_mma();
      WS_STAMP(0, k, 5);           // G2 complete
      {
        float o[16];
        tc::tmem_ld16(tmem_row + SM::t_g1s, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = fmaxf(o[c] + sw[W::v_bf + c], 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) store_chunk(p_s, SM::p_img + q, t, o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
      WS_STAMP(0, k, 6);           // E2 done
      publish();
      // ===== G3: the img columns of lr0 (rows 8..23) and of color.0's shared part (rows 72..87) =====
      if (warp == 0) {
        const uint32_t id = tc::idesc_tf32(128, 64);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          tc::mma_tf32_elect(tmem_u + SM::t_lr0, desc_a(p_addr, SM::p_img + 2 * kk, 2048u), desc_b(W::b0, 2 + 2 * kk, 64), id, 1);
          tc::mma_tf32_elect(tmem_u + SM::t_cs, desc_a(p_addr, SM::p_img + 2 * kk, 2048u), desc_b(W::bc_shared, 18 + 2 * kk, 64), id, 1);
        }
        tc::mma_commit_elect(&mma_bar);
        __syncwarp();
      }
      wait_mma();
      WS_STAMP(0, k, 7);           // G3 complete
      float sigma;
      {
        float sg = sw[W::v_bs];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float x[32];
          tc::tmem_ld32(tmem_row + SM::t_lr0 + half * 32, x);
          tc::tmem_ld_wait();
#pragma unroll
          for (int jx = 0; jx < 32; ++jx) {
            x[jx] = fmaxf(x[jx] + sw[W::v_b0 + half * 32 + jx], 0.f);
            sg = fmaf(x[jx], sw[W::v_ws + half * 32 + jx], sg);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) store_chunk(p_s, SM::p_x + half * 8 + q, t, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        }
        sigma = (sg > 20.f) ? sg : log1pf(expf(sg));
      }
      WS_STAMP(0, k, 8);           // E3 done
      publish();
      // ===== G5: the x columns of color.0's shared part (rows 0..63) =====
      if (warp == 0) {
        const uint32_t id = tc::idesc_tf32(128, 64);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) tc::mma_tf32_elect(tmem_u + SM::t_cs, desc_a(p_addr, SM::p_x + 2 * kk, 2048u), desc_b(W::bc_shared, 2 * kk, 64), id, 1);
        tc::mma_commit_elect(&mma_bar);
        __syncwarp();
      }
      wait_mma();
      WS_STAMP(0, k, 9);           // G5 complete
      // ---- E5: color.0 = ReLU(shared + view), color.2 logits, softmax over views, blend of the SOURCE colours (nerf.py:38-43) ----
      float cr = 0.f, cg = 0.f, cb = 0.f;
      {
        float cl[S];
#pragma unroll
        for (int s = 0; s < S; ++s) cl[s] = sw[W::v_b2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float sh[32];
          tc::tmem_ld32(tmem_row + SM::t_cs + half * 32, sh);
          tc::tmem_ld_wait();
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float h[32];
            tc::tmem_ld32(tmem_row + SM::t_cv + 64 * s + half * 32, h);
            tc::tmem_ld_wait();
#pragma unroll
            for (int jx = 0; jx < 32; ++jx) cl[s] = fmaf(fmaxf(sh[jx] + h[jx], 0.f), sw[W::v_w2 + half * 32 + jx], cl[s]);   // bias in the view accumulator
          }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          cl[s] = fmaxf(cl[s], 0.f);
          mx = fmaxf(mx, cl[s]);
        }
        float den = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          cl[s] = expf(cl[s] - mx);
          den += cl[s];
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float ws_ = cl[s] / den;
          cr = fmaf(rgb_s[s][0], ws_, cr), cg = fmaf(rgb_s[s][1], ws_, cg), cb = fmaf(rgb_s[s][2], ws_, cb);
        }
      }

      WS_STAMP(0, k, 10);          // E5 done
      // ================= raw2outputs: prefix product / sums over the Ns lanes of a ray =================
      if (p.out_raw) {   // layered mode: samples are merged across layers by enerf_composite_layers
        if (valid) {
          const size_t o = (size_t)ray * p.out_stride + p.out_off + ks;
          *reinterpret_cast<float4*>(p.out_raw + o * 4) = make_float4(cr, cg, cb, sigma);
          p.out_z[o] = p.depth_inv ? 1.0f / z : z;
        }
      } else if (Ns == 2) {
        // two samples of a ray in lanes (2i, 2i+1): one xor-shuffle exchange per quantity
        const float alpha = 1.f - expf(-sigma);
        const float tr = 1.f - alpha + 1e-10f;
        const float tr_o = __shfl_xor_sync(0xffffffffu, tr, 1);
        const float wk = alpha * (ks == 0 ? 1.f : tr_o);
        const float wk_o = __shfl_xor_sync(0xffffffffu, wk, 1);
        const float cr_o = __shfl_xor_sync(0xffffffffu, cr, 1), cg_o = __shfl_xor_sync(0xffffffffu, cg, 1),
                    cb_o = __shfl_xor_sync(0xffffffffu, cb, 1), z_o = __shfl_xor_sync(0xffffffffu, z, 1);
        const float mx = fmaxf(wk, wk_o);
        const float e = expf(wk - mx), e_o = expf(wk_o - mx);
        if (valid && ks == 0) {
          const float den = e + e_o;                       // same order as the sequential sum over samples
          const float wn0 = e / den, wn1 = e_o / den;
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
      } else {
        const int gbase = lane - ks;                 // first lane of this ray's group (Ns | 32)
        const float alpha = 1.f - expf(-sigma);
        const float tr = 1.f - alpha + 1e-10f;
        float T = 1.f;
        for (int jx = 0; jx + 1 < Ns; ++jx) {
          const float tj = __shfl_sync(0xffffffffu, tr, gbase + jx);
          if (jx < ks) T *= tj;
        }
        const float wk = alpha * T;
        float ar = 0.f, ag = 0.f, ab = 0.f, mx = -INFINITY;
        for (int jx = 0; jx < Ns; ++jx) {
          const float wj = __shfl_sync(0xffffffffu, wk, gbase + jx);
          ar = fmaf(wj, __shfl_sync(0xffffffffu, cr, gbase + jx), ar);
          ag = fmaf(wj, __shfl_sync(0xffffffffu, cg, gbase + jx), ag);
          ab = fmaf(wj, __shfl_sync(0xffffffffu, cb, gbase + jx), ab);
          mx = fmaxf(mx, wj);
        }
        const float e = expf(wk - mx);
        float den = 0.f;
        for (int jx = 0; jx < Ns; ++jx) den += __shfl_sync(0xffffffffu, e, gbase + jx);
        const float wn = e / den;
        float dsum = 0.f, wsum = 0.f;
        const float wz = wn * z;
        for (int jx = 0; jx < Ns; ++jx) {
          dsum += __shfl_sync(0xffffffffu, wz, gbase + jx);
          wsum += __shfl_sync(0xffffffffu, wn, gbase + jx);
        }
        if (valid) {
          p.out_weights[(size_t)ray * Ns + ks] = wn;
          if (ks == 0) {
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
      WS_STAMP(0, k, 11);          // composited + stored
      // the next tile's first batch overwrites the accumulators: its issue is ordered behind the consumer barrier at the top
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, SM::tmem_cols);
}

template <int S>
static int launch(const Params& P, cudaStream_t stream) {
  constexpr size_t smem = Smem<S>::bytes;
  static PerDeviceSize attr_set;   // the attribute (and the SM count) is per device
  if (attr_set.cur() < smem) {
    cudaError_t e = cudaFuncSetAttribute(render_rays_ws_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("render_rays_ws: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
      return ENERF_ECUDA;
    }
    attr_set.cur() = smem;
  }
  const int n_sm = device_sm_count();
  const int grid = P.n_tiles < n_sm ? P.n_tiles : n_sm;    // persistent: one CTA per SM
  render_rays_ws_kernel<S><<<grid, 384, smem, stream>>>(P);
  ENERF_CHECK_LAUNCH("render_rays_ws");
  return ENERF_OK;
}

}  // namespace ws

// 0 auto, 1 the single-role kernel (render_rays_tc.cuh), 2 this kernel for 2-3 views.  auto = 1: measured on the B200 at the
// headline workload the two kernels are on par (0.316 vs 0.305 ms, profiles/r2_frame_ab.md), so the simpler one stays the default.
static int g_ray_impl = 0;

// Called by enerf_render_rays_tc / enerf_render_rays_raw_tc (render_rays_tc.cu).  Returns 1 when this kernel does not
// take the configuration (the caller then launches render_rays_tc_kernel).
int render_rays_ws_try_launch(const RayParams& r, const float* wblob, int n_tiles, cudaStream_t stream, unsigned long long* dbg) {
  if (g_ray_impl != 2 || r.S < 2 || r.S > 3) return 1;
  ws::Params P;
  P.r = r, P.wblob = wblob, P.n_tiles = n_tiles, P.dbg = dbg;
  return r.S == 2 ? ws::launch<2>(P, stream) : ws::launch<3>(P, stream);
}

}  // namespace enerf

// Diagnostic: impl 0 = auto (the single-role kernel), 1 = the single-role kernel, 2 = the warp-specialised kernel for 2-3 views.
extern "C" int enerf_render_rays_tc_select(int impl) {
  enerf::g_ray_impl = impl;
  return ENERF_OK;
}
