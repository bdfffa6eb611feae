// render_rays.cu -- C-ABI dispatcher of the fused ray stage (kernel in render_rays.cuh; one
// translation unit per (feat_ch, n_views) instantiation, see render_rays_inst.cu, so that the
// heavily unrolled kernels compile in parallel).
#include "common.cuh"

namespace enerf {
struct RayParams;
template <int FC, int SMAX, bool STATIC_S>
int launch_rays(const RayParams& p, cudaStream_t stream);
}  // namespace enerf
#include "render_rays_params.cuh"

static int dispatch_rays(const enerf::RayParams& p, int feat_ch, int n_views, cudaStream_t stream) {
  using namespace enerf;
  if (feat_ch == 8) {
    switch (n_views) {
      case 2: return launch_rays<11, 2, true>(p, stream);
      case 3: return launch_rays<11, 3, true>(p, stream);
      case 4: return launch_rays<11, 4, true>(p, stream);
      default: return launch_rays<11, ENERF_MAX_VIEWS, false>(p, stream);
    }
  }
  switch (n_views) {
    case 2: return launch_rays<35, 2, true>(p, stream);
    case 3: return launch_rays<35, 3, true>(p, stream);
    case 4: return launch_rays<35, 4, true>(p, stream);
    default: return launch_rays<35, ENERF_MAX_VIEWS, false>(p, stream);
  }
}

extern "C" int enerf_render_rays(const EnerfCam* cam, int level, const float* const* weights, int n_weights,
                                 const float* rays, int n_rays, const float* depth, const float* std, const float* near_far,
                                 int hv, int wv, const float* feat_vol, int D, int vol_row0, int vol_rows, const float* img_feat_rgb,
                                 int n_views, int Hr, int Wr, int feat_ch, int num_samples, int depth_inv, int white_bkgd,
                                 int viewdir_agg, const int* n_rays_dev, float* out_rgb, float* out_depth, float* out_weights,
                                 void* stream_) {
  using namespace enerf;
  cudaStream_t stream = (cudaStream_t)stream_;
  ENERF_REQUIRE(cam && weights && rays && depth && std && near_far && feat_vol && img_feat_rgb && out_rgb && out_depth && out_weights,
                ENERF_EINVAL, "render_rays: null pointer");
  ENERF_REQUIRE(n_weights == 16, ENERF_EINVAL, "render_rays: expected 16 weight pointers, got %d", n_weights);
  ENERF_REQUIRE(n_views >= 2 && n_views <= ENERF_MAX_VIEWS, ENERF_EINVAL,
                "render_rays: n_views %d not in [2,%d] (unbiased variance over views needs >= 2, nerf.py:82)", n_views, ENERF_MAX_VIEWS);
  ENERF_REQUIRE(num_samples >= 1 && num_samples <= 8, ENERF_EUNSUPPORTED, "render_rays: num_samples %d not in [1,8]", num_samples);
  ENERF_REQUIRE(level >= 0 && level < ENERF_MAX_LEVELS, ENERF_EINVAL, "render_rays: level %d", level);
  ENERF_REQUIRE(feat_ch == 8 || feat_ch == 32, ENERF_EUNSUPPORTED, "render_rays: feat_ch %d not in {8,32}", feat_ch);
  ENERF_REQUIRE(vol_row0 >= 0 && vol_rows > 0 && vol_row0 + vol_rows <= hv, ENERF_EINVAL, "render_rays: volume rows [%d,%d) outside [0,%d)",
                vol_row0, vol_row0 + vol_rows, hv);
  if (n_rays <= 0) return ENERF_OK;
  RayParams p;
  p.cam = cam, p.level = level;
  for (int i = 0; i < 16; ++i) p.w[i] = weights[i];
  p.rays = rays, p.n_rays = n_rays, p.depth = depth, p.std = std, p.near_far = near_far, p.hv = hv, p.wv = wv;
  p.feat_vol = feat_vol, p.D = D, p.vol_y0 = vol_row0, p.vol_h = vol_rows, p.n_rays_dev = n_rays_dev;
  p.img = img_feat_rgb, p.S = n_views, p.Hr = Hr, p.Wr = Wr;
  p.num_samples = num_samples, p.depth_inv = depth_inv, p.white_bkgd = white_bkgd, p.viewdir_agg = viewdir_agg;
  p.out_rgb = out_rgb, p.out_depth = out_depth, p.out_weights = out_weights;
  p.win_x = p.win_y = p.win_w = 0, p.out_raw = p.out_z = nullptr, p.out_stride = p.out_off = 0;
  return dispatch_rays(p, feat_ch, n_views, stream);
}

extern "C" int enerf_render_rays_raw(const EnerfCam* cam, int level, const float* const* weights, int n_weights,
                                     const float* rays, const int* window, const float* depth, const float* std,
                                     const float* near_far, int hv, int wv, const float* img_feat_rgb, int n_views, int Hr, int Wr,
                                     int feat_ch, int num_samples, int depth_inv, int viewdir_agg, float* out_raw, float* out_z,
                                     int out_stride, int out_off, void* stream_) {
  using namespace enerf;
  cudaStream_t stream = (cudaStream_t)stream_;
  ENERF_REQUIRE(cam && weights && rays && window && depth && std && near_far && img_feat_rgb && out_raw && out_z, ENERF_EINVAL,
                "render_rays_raw: null pointer");
  ENERF_REQUIRE(n_weights == 16, ENERF_EINVAL, "render_rays_raw: expected 16 weight pointers, got %d", n_weights);
  ENERF_REQUIRE(n_views >= 2 && n_views <= ENERF_MAX_VIEWS, ENERF_EINVAL, "render_rays_raw: n_views %d not in [2,%d]", n_views,
                ENERF_MAX_VIEWS);
  ENERF_REQUIRE(num_samples >= 1 && num_samples <= 8, ENERF_EUNSUPPORTED, "render_rays_raw: num_samples %d not in [1,8]", num_samples);
  ENERF_REQUIRE(level >= 0 && level < ENERF_MAX_LEVELS, ENERF_EINVAL, "render_rays_raw: level %d", level);
  ENERF_REQUIRE(feat_ch == 8 || feat_ch == 32, ENERF_EUNSUPPORTED, "render_rays_raw: feat_ch %d not in {8,32}", feat_ch);
  const int x = window[0], y = window[1], w = window[2], h = window[3];
  ENERF_REQUIRE(x >= 0 && y >= 0 && w >= 0 && h >= 0 && x + w <= Wr && y + h <= Hr, ENERF_EINVAL,
                "render_rays_raw: window (%d,%d,%d,%d) outside the %dx%d frame", x, y, w, h, Wr, Hr);
  ENERF_REQUIRE(out_off >= 0 && out_off + num_samples <= out_stride, ENERF_EINVAL, "render_rays_raw: slot [%d,%d) outside stride %d",
                out_off, out_off + num_samples, out_stride);
  if (w == 0 || h == 0) return ENERF_OK;
  RayParams p;
  p.cam = cam, p.level = level;
  for (int i = 0; i < 16; ++i) p.w[i] = weights[i];
  p.rays = rays, p.n_rays = w * h, p.depth = depth, p.std = std, p.near_far = near_far, p.hv = hv, p.wv = wv;
  p.feat_vol = nullptr, p.D = 1, p.vol_y0 = 0, p.vol_h = hv, p.n_rays_dev = nullptr;
  p.img = img_feat_rgb, p.S = n_views, p.Hr = Hr, p.Wr = Wr;
  p.num_samples = num_samples, p.depth_inv = depth_inv, p.white_bkgd = 0, p.viewdir_agg = viewdir_agg;
  p.out_rgb = p.out_depth = p.out_weights = nullptr;
  p.win_x = x, p.win_y = y, p.win_w = w, p.out_raw = out_raw, p.out_z = out_z, p.out_stride = out_stride, p.out_off = out_off;
  return dispatch_rays(p, feat_ch, n_views, stream);
}
