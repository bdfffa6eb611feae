// camera.cu -- per-frame camera derivations on device (no host sync) + library-wide error state.
//
// Replaces get_proj_mats (/root/reference/lib/networks/enerf/utils.py:35-55: K_s*E_s*inv(P_t)) and
// the torch.inverse calls in get_img_feat (utils.py:707-708: camera centres).  The reference runs
// these as fp32 LU factorisations on the GPU with a host sync each; here one tiny kernel does the
// closed-form affine inverses in fp64 and rounds once to fp32.
#include <stdarg.h>

#include "common.cuh"

namespace enerf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// inverse of the 4x4 affine matrix [M t; 0 0 0 1] given as 3x4 row-major (double)
__device__ void affine_inverse(const double* A, double* inv /*3x4*/) {
  const double a = A[0], b = A[1], c = A[2], d = A[4], e = A[5], f = A[6], g = A[8], h = A[9], i = A[10];
  const double co00 = e * i - f * h, co01 = f * g - d * i, co02 = d * h - e * g;
  const double det = a * co00 + b * co01 + c * co02;
  const double r = 1.0 / det;
  double m[9];
  m[0] = co00 * r;
  m[1] = (c * h - b * i) * r;
  m[2] = (b * f - c * e) * r;
  m[3] = co01 * r;
  m[4] = (a * i - c * g) * r;
  m[5] = (c * d - a * f) * r;
  m[6] = co02 * r;
  m[7] = (b * g - a * h) * r;
  m[8] = (a * e - b * d) * r;
  for (int rr = 0; rr < 3; ++rr) {
    inv[rr * 4 + 0] = m[rr * 3 + 0];
    inv[rr * 4 + 1] = m[rr * 3 + 1];
    inv[rr * 4 + 2] = m[rr * 3 + 2];
    inv[rr * 4 + 3] = -(m[rr * 3 + 0] * A[3] + m[rr * 3 + 1] * A[7] + m[rr * 3 + 2] * A[11]);
  }
}

struct Scales {
  float v[ENERF_MAX_LEVELS * 3];
};

__global__ void camera_setup_kernel(const float* __restrict__ src_exts, const float* __restrict__ src_ixts,
                                    const float* __restrict__ tar_ext, const float* __restrict__ tar_ixt,
                                    const float* __restrict__ near_far, int S, int L, Scales sc, EnerfCam* cam) {
  const int t = threadIdx.x;  // one thread per (level, view)
  if (t == 0) {
    double E[12], inv[12];
    for (int k = 0; k < 12; ++k) E[k] = tar_ext[k];
    affine_inverse(E, inv);
    cam->tar_center[0] = (float)inv[3];
    cam->tar_center[1] = (float)inv[7];
    cam->tar_center[2] = (float)inv[11];
    cam->near_far[0] = near_far[0];
    cam->near_far[1] = near_far[1];
  }
  if (t >= S * L) return;
  const int l = t / S, s = t % S;
  const float* Es = src_exts + s * 16;
  const float* Ks = src_ixts + s * 9;
  if (l == 0) {
    double E[12], inv[12];
    for (int k = 0; k < 12; ++k) {
      E[k] = Es[k];
      cam->src_ext[s][k] = Es[k];
    }
    affine_inverse(E, inv);
    cam->src_center[s][0] = (float)inv[3];
    cam->src_center[s][1] = (float)inv[7];
    cam->src_center[s][2] = (float)inv[11];
  }
  const double im_scale = sc.v[l * 3 + 0], vol_scale = sc.v[l * 3 + 1];
  const float render_scale = sc.v[l * 3 + 2];
  for (int k = 0; k < 9; ++k) cam->src_ixt[l][s][k] = (k < 6) ? Ks[k] * render_scale : Ks[k];
  // P_s = (K_s rows 0,1 * im_scale) * E_s[:3]   (3x4)
  double Ps[12], Pt[12], Pti[12];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      double a = 0, b = 0;
      for (int k = 0; k < 3; ++k) {
        const double ks = (double)Ks[r * 3 + k] * (r < 2 ? im_scale : 1.0);
        const double kt = (double)tar_ixt[r * 3 + k] * (r < 2 ? vol_scale : 1.0);
        a += ks * (double)Es[k * 4 + c];
        b += kt * (double)tar_ext[k * 4 + c];
      }
      Ps[r * 4 + c] = a;
      Pt[r * 4 + c] = b;
    }
  affine_inverse(Pt, Pti);
  // homo = P_s (3x4) * [Pti; 0 0 0 1] (4x4) -> 3x4
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      double a = (c == 3) ? Ps[r * 4 + 3] : 0.0;
      for (int k = 0; k < 3; ++k) a += Ps[r * 4 + k] * Pti[k * 4 + c];
      cam->homo[l][s][r * 4 + c] = (float)a;
    }
}

}  // namespace enerf

extern "C" int enerf_abi_version(void) { return ENERF_ABI_VERSION; }
extern "C" const char* enerf_last_error(void) { return enerf::g_err; }

extern "C" int enerf_camera_setup(const float* src_exts, const float* src_ixts, const float* tar_ext,
                                  const float* tar_ixt, const float* near_far, int n_views, int n_levels,
                                  const float* scales, EnerfCam* cam_out, void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(src_exts && src_ixts && tar_ext && tar_ixt && near_far && scales && cam_out, ENERF_EINVAL,
                "camera_setup: null pointer");
  ENERF_REQUIRE(n_views >= 1 && n_views <= ENERF_MAX_VIEWS, ENERF_EINVAL, "camera_setup: n_views %d not in [1,%d]",
                n_views, ENERF_MAX_VIEWS);
  ENERF_REQUIRE(n_levels >= 1 && n_levels <= ENERF_MAX_LEVELS, ENERF_EINVAL, "camera_setup: n_levels %d not in [1,%d]",
                n_levels, ENERF_MAX_LEVELS);
  Scales sc;
  for (int i = 0; i < n_levels * 3; ++i) sc.v[i] = scales[i];
  camera_setup_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(src_exts, src_ixts, tar_ext, tar_ixt, near_far, n_views,
                                                          n_levels, sc, cam_out);
  ENERF_CHECK_LAUNCH("camera_setup");
  return ENERF_OK;
}

// ---- on-device ray generation ---------------------------------------
// Replaces the numpy ray builder of the data layer (/root/reference/lib/datasets/enerf_utils.py:25-32,
// 60-71, 'test' branch): rays[v*W+u] = { c2w[:3,3], [u,v,1] @ inv(K*scale)^T @ c2w[:3,:3]^T, u, v }.
// The per-frame 3x3 (R_c2w * K^-1) is formed in fp64 like numpy does, the per-ray product is fp64
// and rounded once to fp32, so the result matches the reference's float32 cast to <= 1 ulp.  This
// removes the largest per-frame H2D transfer (10.5 MB at 512x640).
namespace enerf {
__global__ void __launch_bounds__(256) generate_rays_kernel(const float* __restrict__ tar_ext, const float* __restrict__ tar_ixt,
                                                            float scale, int W, int row0, int n_rows, float* __restrict__ rays) {
  __shared__ double M[9];
  __shared__ double org[3];
  if (threadIdx.x == 0) {
    double E[12], Ei[12], K[12], Ki[12];
    for (int k = 0; k < 12; ++k) E[k] = tar_ext[k];
    affine_inverse(E, Ei);  // c2w (3x4)
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) K[r * 4 + c] = (double)tar_ixt[r * 3 + c] * (r < 2 ? (double)scale : 1.0);
      K[r * 4 + 3] = 0.0;
    }
    affine_inverse(K, Ki);  // K^-1 in the 3x3 block
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) M[r * 3 + c] = Ei[r * 4 + 0] * Ki[0 * 4 + c] + Ei[r * 4 + 1] * Ki[1 * 4 + c] + Ei[r * 4 + 2] * Ki[2 * 4 + c];
    org[0] = Ei[3], org[1] = Ei[7], org[2] = Ei[11];
  }
  __syncthreads();
  const long long total = (long long)n_rows * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % W), v = row0 + (int)(i / W);
    const double du = u, dv = v;
    float4* o = reinterpret_cast<float4*>(rays + i * 8);
    o[0] = make_float4((float)org[0], (float)org[1], (float)org[2], (float)(M[0] * du + M[1] * dv + M[2]));
    o[1] = make_float4((float)(M[3] * du + M[4] * dv + M[5]), (float)(M[6] * du + M[7] * dv + M[8]), (float)u, (float)v);
  }
}
}  // namespace enerf

extern "C" int enerf_generate_rays(const float* tar_ext, const float* tar_ixt, float scale, int W, int row0, int n_rows, float* rays,
                                   void* stream) {
  using namespace enerf;
  ENERF_REQUIRE(tar_ext && tar_ixt && rays, ENERF_EINVAL, "generate_rays: null pointer");
  ENERF_REQUIRE(W > 0 && n_rows >= 0 && row0 >= 0, ENERF_EINVAL, "generate_rays: bad dims");
  if (n_rows == 0) return ENERF_OK;
  const long long total = (long long)n_rows * W;
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  generate_rays_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tar_ext, tar_ixt, scale, W, row0, n_rows, rays);
  ENERF_CHECK_LAUNCH("generate_rays");
  return ENERF_OK;
}
