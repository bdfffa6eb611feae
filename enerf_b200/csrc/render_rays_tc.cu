// render_rays_tc.cu -- C-ABI entry points of the tensor-core ray stage and the dispatch between its two kernels:
// the warp-specialised kernel (render_rays_ws.cu; 2-3 source views) and the single-role kernel (render_rays_tc.cuh;
// 2-8 views, one instantiation per view count in render_rays_tc_inst.cu).
#include "render_rays_tc.cuh"

namespace enerf {

// render_rays_ws.cu: the warp-specialised kernel (2-3 views); returns 1 when it does not take the configuration
int render_rays_ws_try_launch(const RayParams& r, const float* wblob, int n_tiles, cudaStream_t stream, unsigned long long* dbg);

extern template int launch_rays_tc<2>(const RayTcParams&, cudaStream_t);
extern template int launch_rays_tc<3>(const RayTcParams&, cudaStream_t);
extern template int launch_rays_tc<4>(const RayTcParams&, cudaStream_t);
extern template int launch_rays_tc<5>(const RayTcParams&, cudaStream_t);
extern template int launch_rays_tc<6>(const RayTcParams&, cudaStream_t);
extern template int launch_rays_tc<7>(const RayTcParams&, cudaStream_t);
extern template int launch_rays_tc<8>(const RayTcParams&, cudaStream_t);

static unsigned long long* g_ray_dbg_host = nullptr;

static int dispatch_rays_tc(const RayTcParams& P, int n_views, cudaStream_t stream) {
  switch (n_views) {
    case 2: return launch_rays_tc<2>(P, stream);
    case 3: return launch_rays_tc<3>(P, stream);
    case 4: return launch_rays_tc<4>(P, stream);
    case 5: return launch_rays_tc<5>(P, stream);
    case 6: return launch_rays_tc<6>(P, stream);
    case 7: return launch_rays_tc<7>(P, stream);
    default: return launch_rays_tc<8>(P, stream);
  }
}

}  // namespace enerf

extern "C" int enerf_render_rays_tc(const EnerfCam* cam, int level, const float* wblob, const float* rays, int n_rays,
                                    const float* depth, const float* std, const float* near_far, int hv, int wv,
                                    const float* feat_vol, int D, int vol_row0, int vol_rows, const float* img_feat_rgb, int n_views,
                                    int Hr, int Wr, int feat_ch, int num_samples, int depth_inv, int white_bkgd, int viewdir_agg,
                                    const int* n_rays_dev, float* out_rgb, float* out_depth, float* out_weights, void* stream_) {
  using namespace enerf;
  cudaStream_t stream = (cudaStream_t)stream_;
  ENERF_REQUIRE(cam && wblob && rays && depth && std && near_far && feat_vol && img_feat_rgb && out_rgb && out_depth && out_weights,
                ENERF_EINVAL, "render_rays_tc: null pointer");
  ENERF_REQUIRE(feat_ch == 8, ENERF_EUNSUPPORTED, "render_rays_tc: feat_ch %d (tensor-core kernel is built for 8)", feat_ch);
  ENERF_REQUIRE(n_views >= 2 && n_views <= ENERF_MAX_VIEWS, ENERF_EUNSUPPORTED, "render_rays_tc: n_views %d not in [2,8]", n_views);
  ENERF_REQUIRE(num_samples == 1 || num_samples == 2 || num_samples == 4 || num_samples == 8, ENERF_EUNSUPPORTED,
                "render_rays_tc: num_samples %d not in {1,2,4,8}", num_samples);
  ENERF_REQUIRE((long long)n_rays * num_samples + 128 < (1ll << 31), ENERF_EUNSUPPORTED, "render_rays_tc: %d rays x %d samples exceed the 32-bit index range",
                n_rays, num_samples);
  ENERF_REQUIRE(level >= 0 && level < ENERF_MAX_LEVELS, ENERF_EINVAL, "render_rays_tc: level %d", level);
  ENERF_REQUIRE(vol_row0 >= 0 && vol_rows > 0 && vol_row0 + vol_rows <= hv, ENERF_EINVAL, "render_rays_tc: volume rows [%d,%d) outside [0,%d)",
                vol_row0, vol_row0 + vol_rows, hv);
  if (n_rays <= 0) return ENERF_OK;
  RayTcParams P;
  RayParams& p = P.r;
  p.cam = cam, p.level = level;
  for (int i = 0; i < 18; ++i) p.w[i] = nullptr;
  p.rays = rays, p.n_rays = n_rays, p.depth = depth, p.std = std, p.near_far = near_far, p.hv = hv, p.wv = wv;
  p.feat_vol = feat_vol, p.D = D, p.vol_y0 = vol_row0, p.vol_h = vol_rows, p.n_rays_dev = n_rays_dev;
  p.img = img_feat_rgb, p.S = n_views, p.Hr = Hr, p.Wr = Wr;
  p.num_samples = num_samples, p.depth_inv = depth_inv, p.white_bkgd = white_bkgd, p.viewdir_agg = viewdir_agg;
  p.out_rgb = out_rgb, p.out_depth = out_depth, p.out_weights = out_weights;
  p.win_x = p.win_y = p.win_w = 0, p.out_raw = p.out_z = nullptr, p.out_stride = p.out_off = 0;
  P.wblob = wblob, P.dbg = g_ray_dbg_host;
  P.n_tiles = (int)(((long long)n_rays * num_samples + 127) / 128);
  {
    const int rc = render_rays_ws_try_launch(P.r, wblob, P.n_tiles, stream, g_ray_dbg_host);
    if (rc != 1) return rc;
  }
  return dispatch_rays_tc(P, n_views, stream);
}

extern "C" int enerf_render_rays_raw_tc(const EnerfCam* cam, int level, const float* wblob, const float* rays, const int* window,
                                        const float* depth, const float* std, const float* near_far, int hv, int wv,
                                        const float* img_feat_rgb, int n_views, int Hr, int Wr, int feat_ch, int num_samples,
                                        int depth_inv, int viewdir_agg, float* out_raw, float* out_z, int out_stride, int out_off,
                                        void* stream_) {
  using namespace enerf;
  cudaStream_t stream = (cudaStream_t)stream_;
  ENERF_REQUIRE(cam && wblob && rays && window && depth && std && near_far && img_feat_rgb && out_raw && out_z, ENERF_EINVAL,
                "render_rays_raw_tc: null pointer");
  ENERF_REQUIRE(feat_ch == 8, ENERF_EUNSUPPORTED, "render_rays_raw_tc: feat_ch %d (tensor-core kernel is built for 8)", feat_ch);
  ENERF_REQUIRE(n_views >= 2 && n_views <= ENERF_MAX_VIEWS, ENERF_EUNSUPPORTED, "render_rays_raw_tc: n_views %d not in [2,8]", n_views);
  ENERF_REQUIRE(num_samples == 1 || num_samples == 2 || num_samples == 4 || num_samples == 8, ENERF_EUNSUPPORTED,
                "render_rays_raw_tc: num_samples %d not in {1,2,4,8}", num_samples);
  ENERF_REQUIRE(level >= 0 && level < ENERF_MAX_LEVELS, ENERF_EINVAL, "render_rays_raw_tc: level %d", level);
  const int x = window[0], y = window[1], w = window[2], h = window[3];
  ENERF_REQUIRE(x >= 0 && y >= 0 && w >= 0 && h >= 0 && x + w <= Wr && y + h <= Hr, ENERF_EINVAL,
                "render_rays_raw_tc: window (%d,%d,%d,%d) outside the %dx%d frame", x, y, w, h, Wr, Hr);
  ENERF_REQUIRE(out_off >= 0 && out_off + num_samples <= out_stride, ENERF_EINVAL, "render_rays_raw_tc: slot [%d,%d) outside stride %d",
                out_off, out_off + num_samples, out_stride);
  if (w == 0 || h == 0) return ENERF_OK;
  ENERF_REQUIRE((long long)w * h * num_samples + 128 < (1ll << 31), ENERF_EUNSUPPORTED, "render_rays_raw_tc: window x samples exceed the 32-bit index range");
  RayTcParams P;
  RayParams& p = P.r;
  p.cam = cam, p.level = level;
  for (int i = 0; i < 18; ++i) p.w[i] = nullptr;
  p.rays = rays, p.n_rays = w * h, p.depth = depth, p.std = std, p.near_far = near_far, p.hv = hv, p.wv = wv;
  p.feat_vol = nullptr, p.D = 1, p.vol_y0 = 0, p.vol_h = hv, p.n_rays_dev = nullptr;
  p.img = img_feat_rgb, p.S = n_views, p.Hr = Hr, p.Wr = Wr;
  p.num_samples = num_samples, p.depth_inv = depth_inv, p.white_bkgd = 0, p.viewdir_agg = viewdir_agg;
  p.out_rgb = p.out_depth = p.out_weights = nullptr;
  p.win_x = x, p.win_y = y, p.win_w = w, p.out_raw = out_raw, p.out_z = out_z, p.out_stride = out_stride, p.out_off = out_off;
  P.wblob = wblob, P.dbg = g_ray_dbg_host;
  P.n_tiles = (int)(((long long)p.n_rays * num_samples + 127) / 128);
  {
    const int rc = render_rays_ws_try_launch(P.r, wblob, P.n_tiles, stream, g_ray_dbg_host);
    if (rc != 1) return rc;
  }
  return dispatch_rays_tc(P, n_views, stream);
}

extern "C" int enerf_render_rays_debug(unsigned long long* buf) {
  enerf::g_ray_dbg_host = buf;     // passed to later launches of the single-role kernel as a launch parameter
  return ENERF_OK;
}
