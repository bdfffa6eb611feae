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
  const float* img;
  int S, Hr, Wr;
  int num_samples, depth_inv, white_bkgd, viewdir_agg;
  float *out_rgb, *out_depth, *out_weights;
};

}  // namespace enerf
