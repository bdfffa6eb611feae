// render_rays_params.cuh -- launch parameters of the fused ray kernel
#pragma once
#include "common.cuh"

namespace enerf {

struct RayParams {
  const EnerfCam* cam;
  int level;
  const float* w[18];
  const float* rays;
  int n_rays;
  const float *depth, *std, *near_far;
  int hv, wv;
  const float* feat_vol;
  int D;
  // rows [vol_y0, vol_y0 + vol_h) of the (D,hv,wv,8) feature volume are resident (feat_vol = that crop, (D,vol_h,wv,8));
  // the full grid is vol_y0 = 0, vol_h = hv.  Used by the row-band multi-GPU layout (enerf_b200/dist.py).
  int vol_y0, vol_h;
  // optional device-side ray count: the launch covers n_rays (an upper bound), rays >= *n_rays_dev are skipped
  const int* n_rays_dev;
  const float* img;
  int S, Hr, Wr;
  int num_samples, depth_inv, white_bkgd, viewdir_agg;
  float *out_rgb, *out_depth, *out_weights;
  // layered ("composite") mode, out_raw != nullptr: rays is the full (Hr*Wr, 8) frame, ray r of the
  // launch is pixel (win_x + r % win_w, win_y + r / win_w), and instead of compositing, the per-sample
  // (r,g,b,sigma) and metric z go to out_raw[(pixel*out_stride + out_off + k)*4] / out_z[pixel*out_stride + out_off + k]
  int win_x, win_y, win_w;
  float *out_raw, *out_z;
  int out_stride, out_off;
};

}  // namespace enerf
