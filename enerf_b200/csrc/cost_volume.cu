// cost_volume.cu -- depth hypotheses, fused homography-warp + variance cost volume, depth regression.
//
//  enerf_depth_hypotheses : get_depth_values      /root/reference/lib/networks/enerf/utils.py:98-151
//  enerf_cost_volume      : build_feature_volume  utils.py:322-349  (homo_warp :57-95 x S, variance :337-345)
//  enerf_depth_regress    : depth_regression      utils.py:658-663
//
// The cost volume is a pure gather: per target voxel, S bilinear 4-tap reads of a C-channel source
// pixel (zeros padding, align_corners=True).  Channels-last sources make every tap one contiguous
// C*4-byte read; C/4 adjacent lanes share a voxel and read the float4 slices of the same tap, so a
// warp issues full 128-byte (C=32) or 64-byte (C=16) segments.  Sum and sum of squares live in
// registers; the S warped volumes and the sampling grid of the reference are never materialised.
// Roofline: L2/HBM (AI ~ 1 FLOP/B); algorithmic bytes per launch = S*hs*ws*C*4 (read once) +
// D*h*w*C*4 (written once).
#include "common.cuh"

namespace enerf {

// plane i of D between end points (a,b): linear in disparity when depth_inv, else linear in depth
// (utils.py:104-111 level 0, :135-146 level >0)
__device__ __forceinline__ float plane_depth(float a, float b, int i, int D, bool depth_inv) {
  const float t = linspace01(i, D);
  if (depth_inv) {
    const float ia = 1.0f / a, ib = 1.0f / b;
    return 1.0f / (ia + t * (ib - ia));
  }
  return a + t * (b - a);
}

__global__ void depth_hypotheses_kernel(const EnerfCam* __restrict__ cam, const float* __restrict__ prev_depth,
                                        const float* __restrict__ prev_std, const float* __restrict__ prev_nf, int hp,
                                        int wp, int h, int w, int D, int depth_inv, float* __restrict__ ends,
                                        float* __restrict__ nf_out, const float* __restrict__ first_nf) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= h * w) return;
  const int x = pix % w, y = pix / w;
  float a, b;
  if (prev_depth == nullptr) {
    a = first_nf ? first_nf[0] : cam->near_far[0];
    b = first_nf ? first_nf[1] : cam->near_far[1];
  } else {
    const float d = bilinear_ac(prev_depth, hp, wp, h, w, y, x);
    const float s = bilinear_ac(prev_std, hp, wp, h, w, y, x);
    const float n0 = bilinear_ac(prev_nf, hp, wp, h, w, y, x);
    const float n1 = bilinear_ac(prev_nf + (size_t)hp * wp, hp, wp, h, w, y, x);
    // previous level lives in disparity space: [d+s, d-s] clamped into its near_far, then inverted
    const float lo = fminf(d + s, n0), hi = fmaxf(d - s, n1);
    a = 1.0f / lo;
    b = 1.0f / hi;
  }
  ends[pix] = a;
  ends[(size_t)h * w + pix] = b;
  float v0 = plane_depth(a, b, 0, D, depth_inv), v1 = plane_depth(a, b, D - 1, D, depth_inv);
  if (depth_inv) {
    v0 = 1.0f / fmaxf(v0, 1e-6f);
    v1 = 1.0f / fmaxf(v1, 1e-6f);
  }
  nf_out[pix] = v0;
  nf_out[(size_t)h * w + pix] = v1;
}

// thread = (voxel, C/SPLIT channels): the homography, the perspective divide and the bilinear weights
// are computed once per view and applied to the thread's float4 slices of the four taps; the SPLIT
// lanes of a voxel read adjacent slices of the same source pixel record.
template <int C, int SPLIT>
__global__ void __launch_bounds__(256) cost_volume_kernel(const EnerfCam* __restrict__ cam, int level,
                                                          const float* __restrict__ feat, int S, int hs, int ws,
                                                          const float* __restrict__ ends, int D, int h, int w,
                                                          int depth_inv, float* __restrict__ var_out, int x0, int y0, int wc,
                                                          int hc, FastDiv div_wc, FastDiv div_hc) {
  // SPLIT lanes share a voxel, each owning C/SPLIT channels of every tap
  constexpr int CH = C / SPLIT;                    // channels per thread
  constexpr int NV = CH / 4;                       // float4 per tap per thread
  __shared__ float Hm[ENERF_MAX_VIEWS * 12];
  for (int e = threadIdx.x; e < S * 12; e += blockDim.x) Hm[e] = cam->homo[level][e / 12][e % 12];
  __syncthreads();
  // the volume covers the window [x0,x0+wc) x [y0,y0+hc) of the h x w target grid (the whole grid
  // for enerf_cost_volume; a layer's bbox for enerf_cost_volume_window, utils.py:284-289)
  const unsigned total = (unsigned)D * hc * wc * SPLIT;      // 32-bit index math (the launcher checks the range)
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int g = (int)(t % SPLIT);
  const unsigned vox = t / SPLIT;
  const unsigned vrow = div_wc.div(vox);            // division by a run-time extent as multiply-high + shift (host-made constants)
  const unsigned dd = div_hc.div(vrow);
  const int x = x0 + (int)(vox - vrow * wc), y = y0 + (int)(vrow - dd * hc), d = (int)dd;
  const int pix = y * w + x;
  const float depth = plane_depth(__ldg(ends + pix), __ldg(ends + (size_t)h * w + pix), d, D, depth_inv);
  const float inv_depth = 1.0f / depth;
  const float fx = (float)x, fy = (float)y;
  float4 s1[NV], s2[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) s1[q] = s2[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < S; ++s) {
    const float* M = Hm + s * 12;
    // p = R (x,y,1)^T + T / depth  (utils.py:72); T/depth as T * (1/depth): <= 1 ulp from the reference
    const float q0 = (M[0] * fx + M[1] * fy + M[2]) + M[3] * inv_depth;
    const float q1 = (M[4] * fx + M[5] * fy + M[6]) + M[7] * inv_depth;
    const float q2 = (M[8] * fx + M[9] * fy + M[10]) + M[11] * inv_depth;
    const float rz = 1.0f / fmaxf(q2, 1e-6f);
    // the reference normalises to [-1,1] (utils.py:83-84) and grid_sample(align_corners=True) maps
    // straight back: the sample position is the source pixel coordinate itself
    const float ix = q0 * rz, iy = q1 * rz;
    float4 v[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ix > -1.0f && ix < (float)ws && iy > -1.0f && iy < (float)hs) {  // false for NaN
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      const int x0 = (int)fx0, y0 = (int)fy0;
      const float txr = (fx0 + 1.f) - ix, txl = ix - fx0, tyb = (fy0 + 1.f) - iy, tyt = iy - fy0;
      // zeros padding without branches: a tap outside the image gets weight 0 and a clamped (valid) address -- it adds +0,
      // exactly what skipping it did; the four loads of a view are unconditional and independent
      const bool okx0 = x0 >= 0, okx1 = x0 + 1 < ws, oky0 = y0 >= 0, oky1 = y0 + 1 < hs;
      const float wgt[4] = {(okx0 && oky0) ? txr * tyb : 0.f, (okx1 && oky0) ? txl * tyb : 0.f, (okx0 && oky1) ? txr * tyt : 0.f,
                            (okx1 && oky1) ? txl * tyt : 0.f};
      const int xa = max(x0, 0), xb = min(x0 + 1, ws - 1), ya = max(y0, 0), yb = min(y0 + 1, hs - 1);
      const float* base = feat + ((size_t)s * hs * ws) * C + g * CH;          // (32-bit offsets inside a view: hs * ws * C < 2^31, checked by the launcher)
      const float* pt[4] = {base + (unsigned)((ya * ws + xa) * C), base + (unsigned)((ya * ws + xb) * C), base + (unsigned)((yb * ws + xa) * C),
                            base + (unsigned)((yb * ws + xb) * C)};
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {
        const float wv = wgt[tp];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          const float4 a = ldg4(pt[tp] + 4 * q);
          v[q].x = fmaf(a.x, wv, v[q].x), v[q].y = fmaf(a.y, wv, v[q].y), v[q].z = fmaf(a.z, wv, v[q].z), v[q].w = fmaf(a.w, wv, v[q].w);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      s1[q].x += v[q].x, s1[q].y += v[q].y, s1[q].z += v[q].z, s1[q].w += v[q].w;
      s2[q].x += v[q].x * v[q].x, s2[q].y += v[q].y * v[q].y, s2[q].z += v[q].z * v[q].z, s2[q].w += v[q].w * v[q].w;
    }
  }
  // var = E[x^2] - E[x]^2 (utils.py:341-345); the 2 C/SPLIT divisions by S as multiplications by 1/S (<= 1 ulp per term; eight IEEE
  // divisions were a sixth of this issue-bound kernel's instructions)
  const float invS = 1.0f / (float)S;
  float4* o = reinterpret_cast<float4*>(var_out + (size_t)vox * C + g * CH);
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    float4 r;
    float m;
    m = s1[q].x * invS, r.x = s2[q].x * invS - m * m;
    m = s1[q].y * invS, r.y = s2[q].y * invS - m * m;
    m = s1[q].z * invS, r.z = s2[q].z * invS - m * m;
    m = s1[q].w * invS, r.w = s2[q].w * invS - m * m;
    o[q] = r;
  }
}

// one thread per pixel: the large maps (level 1: 81,920 pixels x 8 planes)
__global__ void depth_regress_px_kernel(const float* __restrict__ prob, const float* __restrict__ ends, int D, int h, int w,
                                     int depth_inv, float* __restrict__ depth_out, float* __restrict__ std_out,
                                     float* __restrict__ mvs_out, int x0, int y0, int wc, int hc) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= h * w) return;
  const float a = ends[pix], b = ends[(size_t)h * w + pix];
  // prob covers the window [x0,x0+wc) x [y0,y0+hc) and is zero-padded to the h x w grid
  // (network_composite.py:102: F.pad(depth_prob)); outside, softmax of zeros = uniform
  const int lx = pix % w - x0, ly = pix / w - y0;
  const bool inside = lx >= 0 && lx < wc && ly >= 0 && ly < hc;
  const int hw = wc * hc;
  prob += inside ? ly * wc + lx : 0;
#define PROB_AT(d_) (inside ? __ldg(prob + (size_t)(d_) * hw) : 0.f)
  float mx = -INFINITY;
  for (int d = 0; d < D; ++d) mx = fmaxf(mx, PROB_AT(d));
  float den = 0.f;
  for (int d = 0; d < D; ++d) den += expf(PROB_AT(d) - mx);
  float mean = 0.f;
  for (int d = 0; d < D; ++d) {
    const float p = expf(PROB_AT(d) - mx) / den;
    float v = plane_depth(a, b, d, D, depth_inv);
    if (depth_inv) v = 1.0f / fmaxf(v, 1e-6f);
    mean += p * v;
  }
  float var = 0.f;
  for (int d = 0; d < D; ++d) {
    const float p = expf(PROB_AT(d) - mx) / den;
    float v = plane_depth(a, b, d, D, depth_inv);
    if (depth_inv) v = 1.0f / fmaxf(v, 1e-6f);
    const float e = v - mean;
    var += p * (e * e);
  }
  depth_out[pix] = mean;
  std_out[pix] = sqrtf(fmaxf(var, 1e-10f));
  if (mvs_out) mvs_out[pix] = depth_inv ? 1.0f / mean : mean;
#undef PROB_AT
}

// 8 lanes share a pixel and split the D planes (lane j takes d = j, j+8, ...): the level-0 map has only
// 5,120 pixels, one thread per pixel left 128 of the 148 SMs idle for 25 us on the frame's critical path.
// softmax over D, E[v], std: four width-8 shuffle reductions (max, sum, mean, variance).
template <int MAXP>   // planes per lane held in registers: D <= 8 * MAXP
__global__ void __launch_bounds__(256) depth_regress_kernel(const float* __restrict__ prob, const float* __restrict__ ends, int D,
                                                            int h, int w, int depth_inv, float* __restrict__ depth_out,
                                                            float* __restrict__ std_out, float* __restrict__ mvs_out, int x0, int y0,
                                                            int wc, int hc) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int sub = gid & 7;
  const int pix = min(gid >> 3, h * w - 1);       // surplus groups of the last block redo the last pixel (all lanes stay in the shuffles)
  const float a = ends[pix], b = ends[(size_t)h * w + pix];
  // prob covers the window [x0,x0+wc) x [y0,y0+hc) and is zero-padded to the h x w grid
  // (network_composite.py:102: F.pad(depth_prob)); outside, softmax of zeros = uniform
  const int lx = pix % w - x0, ly = pix / w - y0;
  const bool inside = lx >= 0 && lx < wc && ly >= 0 && ly < hc;
  const int hw = wc * hc;
  prob += inside ? ly * wc + lx : 0;
  auto red_max = [](float v) {
    for (int o = 4; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  };
  auto red_sum = [](float v) {
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  float pv[MAXP], vv[MAXP];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int d = sub + 8 * i;
    pv[i] = (d < D) ? (inside ? __ldg(prob + (size_t)d * hw) : 0.f) : -INFINITY;
    mx = fmaxf(mx, pv[i]);
  }
  mx = red_max(mx);
  float den = 0.f;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int d = sub + 8 * i;
    pv[i] = (d < D) ? expf(pv[i] - mx) : 0.f;
    den += pv[i];
    float v = plane_depth(a, b, min(d, D - 1), D, depth_inv);
    if (depth_inv) v = 1.0f / fmaxf(v, 1e-6f);
    vv[i] = v;
  }
  den = red_sum(den);
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    pv[i] = pv[i] / den;
    mean += pv[i] * vv[i];
  }
  mean = red_sum(mean);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const float e = vv[i] - mean;
    var += pv[i] * (e * e);
  }
  var = red_sum(var);
  if (sub == 0 && (gid >> 3) < h * w) {
    depth_out[pix] = mean;
    std_out[pix] = sqrtf(fmaxf(var, 1e-10f));
    if (mvs_out) mvs_out[pix] = depth_inv ? 1.0f / mean : mean;
  }
}

}  // namespace enerf

extern "C" int enerf_depth_hypotheses(const EnerfCam* cam, const float* prev_depth, const float* prev_std,
                                      const float* prev_near_far, int hp, int wp, int h, int w, int D, int depth_inv,
                                      float* ends, float* near_far_out, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(cam && ends && near_far_out, ENERF_EINVAL, "depth_hypotheses: null pointer");
  ENERF_REQUIRE(h > 0 && w > 0 && D >= 1, ENERF_EINVAL, "depth_hypotheses: bad dims h=%d w=%d D=%d", h, w, D);
  if (prev_depth) ENERF_REQUIRE(prev_std && prev_near_far && hp > 0 && wp > 0, ENERF_EINVAL, "depth_hypotheses: prev level incomplete");
  depth_hypotheses_kernel<<<ceil_div(h * w, 256), 256, 0, (cudaStream_t)stream>>>(cam, prev_depth, prev_std, prev_near_far,
                                                                                 hp, wp, h, w, D, depth_inv, ends, near_far_out, nullptr);
  ENERF_CHECK_LAUNCH("depth_hypotheses");
  return ENERF_OK;
}

extern "C" int enerf_depth_hypotheses_layer(const float* layer_near_far, const float* prev_depth, const float* prev_std,
                                            const float* prev_near_far, int hp, int wp, int h, int w, int D, int depth_inv,
                                            float* ends, float* near_far_out, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(ends && near_far_out, ENERF_EINVAL, "depth_hypotheses_layer: null pointer");
  ENERF_REQUIRE(prev_depth || layer_near_far, ENERF_EINVAL, "depth_hypotheses_layer: need the layer's near/far or a previous level");
  ENERF_REQUIRE(h > 0 && w > 0 && D >= 1, ENERF_EINVAL, "depth_hypotheses_layer: bad dims h=%d w=%d D=%d", h, w, D);
  if (prev_depth) ENERF_REQUIRE(prev_std && prev_near_far && hp > 0 && wp > 0, ENERF_EINVAL, "depth_hypotheses_layer: prev level incomplete");
  depth_hypotheses_kernel<<<ceil_div(h * w, 256), 256, 0, (cudaStream_t)stream>>>(nullptr, prev_depth, prev_std, prev_near_far, hp, wp,
                                                                                 h, w, D, depth_inv, ends, near_far_out, layer_near_far);
  ENERF_CHECK_LAUNCH("depth_hypotheses_layer");
  return ENERF_OK;
}

static int cost_volume_launch(const EnerfCam* cam, int level, const float* feat, int S, int C, int hs, int ws, const float* ends, int D,
                              int h, int w, int depth_inv, float* variance, int x0, int y0, int wc, int hc, void* stream_);

extern "C" int enerf_cost_volume(const EnerfCam* cam, int level, const float* feat, int S, int C, int hs, int ws,
                                 const float* ends, int D, int h, int w, int depth_inv, float* variance, void* stream_) {
  return cost_volume_launch(cam, level, feat, S, C, hs, ws, ends, D, h, w, depth_inv, variance, 0, 0, w, h, stream_);
}

extern "C" int enerf_cost_volume_window(const EnerfCam* cam, int level, const float* feat, int S, int C, int hs, int ws,
                                        const float* ends, int D, int h, int w, const int* window, int depth_inv, float* variance,
                                        void* stream_) {
  using namespace enerf;
  ENERF_REQUIRE(window, ENERF_EINVAL, "cost_volume_window: null window");
  const int x0 = window[0], y0 = window[1], wc = window[2], hc = window[3];
  ENERF_REQUIRE(x0 >= 0 && y0 >= 0 && wc > 0 && hc > 0 && x0 + wc <= w && y0 + hc <= h, ENERF_EINVAL,
                "cost_volume_window: window (%d,%d,%d,%d) outside the %dx%d grid", x0, y0, wc, hc, w, h);
  return cost_volume_launch(cam, level, feat, S, C, hs, ws, ends, D, h, w, depth_inv, variance, x0, y0, wc, hc, stream_);
}

static int cost_volume_launch(const EnerfCam* cam, int level, const float* feat, int S, int C, int hs, int ws, const float* ends, int D,
                              int h, int w, int depth_inv, float* variance, int x0, int y0, int wc, int hc, void* stream_) {
  using namespace enerf;
  cudaStream_t stream = (cudaStream_t)stream_;
  ENERF_REQUIRE(cam && feat && ends && variance, ENERF_EINVAL, "cost_volume: null pointer");
  ENERF_REQUIRE(level >= 0 && level < ENERF_MAX_LEVELS, ENERF_EINVAL, "cost_volume: level %d", level);
  ENERF_REQUIRE(S >= 1 && S <= ENERF_MAX_VIEWS, ENERF_EINVAL, "cost_volume: n_views %d not in [1,%d]", S, ENERF_MAX_VIEWS);
  // lanes per voxel (measured A/B at the headline sizes): C = 32: 8 lanes, one float4 of every tap each, so a
  // warp instruction reads whole 128 B source records (67 -> 54 us); C = 16: 2 lanes (4 lanes: 53 -> 68 us)
  const int split = (C == 32) ? 8 : (C >= 16 ? 2 : 1);
  const long long total = (long long)D * hc * wc * split;
  ENERF_REQUIRE(total < (1ll << 31), ENERF_EUNSUPPORTED, "cost_volume: %lld work items exceed the 32-bit index range", total);
  ENERF_REQUIRE((long long)hs * ws * C < (1ll << 31), ENERF_EUNSUPPORTED, "cost_volume: a %dx%dx%d source view exceeds the 32-bit offset range", hs, ws, C);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  const FastDiv div_wc = FastDiv::make((unsigned)wc), div_hc = FastDiv::make((unsigned)hc);
#define CV_LAUNCH(C_, SP_) \
  cost_volume_kernel<C_, SP_><<<blocks, 256, 0, stream>>>(cam, level, feat, S, hs, ws, ends, D, h, w, depth_inv, variance, x0, y0, wc, hc, div_wc, div_hc)
  if (C == 8 && split == 2) CV_LAUNCH(8, 2);
  else if (C == 8) CV_LAUNCH(8, 1);
  else if (C == 16 && split == 4) CV_LAUNCH(16, 4);
  else if (C == 16) CV_LAUNCH(16, 2);
  else if (C == 32 && split == 8) CV_LAUNCH(32, 8);
  else if (C == 32) CV_LAUNCH(32, 2);
  else ENERF_REQUIRE(false, ENERF_EUNSUPPORTED, "cost_volume: C=%d not in {8,16,32}", C);
#undef CV_LAUNCH
  ENERF_CHECK_LAUNCH("cost_volume");
  return ENERF_OK;
}

extern "C" int enerf_depth_regress(const float* depth_prob, const float* ends, int D, int h, int w, int depth_inv,
                                   float* depth, float* std, float* depth_mvs, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(depth_prob && ends && depth && std, ENERF_EINVAL, "depth_regress: null pointer");
  ENERF_REQUIRE(D >= 1, ENERF_EINVAL, "depth_regress: D=%d", D);
  if (h * w >= 148 * 256 || D > 128)     // large maps, or more planes than the 8-lane kernel holds in registers
    depth_regress_px_kernel<<<ceil_div(h * w, 256), 256, 0, (cudaStream_t)stream>>>(depth_prob, ends, D, h, w, depth_inv, depth, std,
                                                                                   depth_mvs, 0, 0, w, h);
  else if (D <= 64)
    depth_regress_kernel<8><<<ceil_div(h * w * 8, 256), 256, 0, (cudaStream_t)stream>>>(depth_prob, ends, D, h, w, depth_inv, depth, std,
                                                                                       depth_mvs, 0, 0, w, h);
  else
    depth_regress_kernel<16><<<ceil_div(h * w * 8, 256), 256, 0, (cudaStream_t)stream>>>(depth_prob, ends, D, h, w, depth_inv, depth, std,
                                                                                        depth_mvs, 0, 0, w, h);
  ENERF_CHECK_LAUNCH("depth_regress");
  return ENERF_OK;
}

extern "C" int enerf_depth_regress_window(const float* depth_prob, const int* window, const float* ends, int D, int h, int w,
                                          int depth_inv, float* depth, float* std, float* depth_mvs, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(depth_prob && window && ends && depth && std, ENERF_EINVAL, "depth_regress_window: null pointer");
  ENERF_REQUIRE(D >= 1, ENERF_EINVAL, "depth_regress_window: D=%d", D);
  const int x0 = window[0], y0 = window[1], wc = window[2], hc = window[3];
  ENERF_REQUIRE(x0 >= 0 && y0 >= 0 && wc > 0 && hc > 0 && x0 + wc <= w && y0 + hc <= h, ENERF_EINVAL,
                "depth_regress_window: window (%d,%d,%d,%d) outside the %dx%d grid", x0, y0, wc, hc, w, h);
  if (h * w >= 148 * 256 || D > 128)
    depth_regress_px_kernel<<<ceil_div(h * w, 256), 256, 0, (cudaStream_t)stream>>>(depth_prob, ends, D, h, w, depth_inv, depth, std,
                                                                                   depth_mvs, x0, y0, wc, hc);
  else if (D <= 64)
    depth_regress_kernel<8><<<ceil_div(h * w * 8, 256), 256, 0, (cudaStream_t)stream>>>(depth_prob, ends, D, h, w, depth_inv, depth, std,
                                                                                       depth_mvs, x0, y0, wc, hc);
  else
    depth_regress_kernel<16><<<ceil_div(h * w * 8, 256), 256, 0, (cudaStream_t)stream>>>(depth_prob, ends, D, h, w, depth_inv, depth, std,
                                                                                        depth_mvs, x0, y0, wc, hc);
  ENERF_CHECK_LAUNCH("depth_regress_window");
  return ENERF_OK;
}

// Test hook (CPU): n / d through the same FastDiv constants the kernels above use for their run-time extents.
extern "C" unsigned enerf_fastdiv_check(unsigned d, unsigned n) { return enerf::FastDiv::make(d).div(n); }
