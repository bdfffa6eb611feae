/* enerf_b200.h -- C ABI of libenerf_b200.so: the B200 (sm_100a) render-time hot path of ENeRF.
 *
 * The reference (zju3dv/ENeRF) is pure Python/PyTorch and has NO FFI; its seam for this path is the
 * Python plugin `lib/networks/make_network.py:5-9` -> `Network.forward(batch)`
 * (`lib/networks/enerf/network.py:76-113`).  This header is the new native seam underneath that
 * plugin: one entry point per stage of `Network.forward`, each citing the reference lines it
 * replaces.  The Python mirror of the plugin (enerf_b200/network.py) binds these with ctypes; the
 * stub a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked [host].
 *  - all tensors are fp32, dense.  Activations are channels-last: images (N,H,W,C), volumes
 *    (D,H,W,C).  The batch inputs keep the reference's layouts (src_inps NCHW, rays (N,8), ...).
 *  - batch size is 1 (the reference's inference loop, run.py:57-76); callers loop for B > 1.
 *  - no allocation, no host synchronisation, no exceptions inside: work is enqueued on `stream`
 *    (a cudaStream_t passed as void*), memory is caller-owned, workspace sizes come from the
 *    *_workspace_bytes functions.
 *  - return 0 on success, a negative ENERF_E* code otherwise; enerf_last_error() gives the text.
 *  - weights are passed as a [host] array of device pointers in the order documented per stage;
 *    they are BN-folded / re-laid-out copies produced by enerf_b200/packing.py.
 */
#ifndef ENERF_B200_H
#define ENERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ENERF_API __attribute__((visibility("default")))
#else
#define ENERF_API
#endif

#define ENERF_ABI_VERSION 2
#define ENERF_MAX_VIEWS 8
#define ENERF_MAX_LEVELS 4

#define ENERF_OK 0
#define ENERF_EINVAL (-1) /* bad argument (shape not divisible, S out of range, null pointer) */
#define ENERF_ECUDA (-2)  /* a CUDA runtime call / kernel launch failed                          */
#define ENERF_EWORKSPACE (-3) /* workspace too small                                             */
#define ENERF_EUNSUPPORTED (-4) /* configuration outside what the kernels were built for         */

/* Per-frame camera quantities, derived ON DEVICE from the batch tensors by enerf_camera_setup so
 * that forward() never synchronises with the host (the reference calls torch.inverse 8x per frame,
 * utils.py:49,707,708).  Lives in device memory; 4-byte floats, no padding surprises. */
typedef struct EnerfCam {
  /* homo[l][s] = row-major 3x4  (K_s * diag(im_feat_scale_l,..,1)) E_s[:3] * inv([K_t*volume_scale_l E_t[:3]; 0 0 0 1])
   * -- get_proj_mats, utils.py:35-55 */
  float homo[ENERF_MAX_LEVELS][ENERF_MAX_VIEWS][12];
  float src_ext[ENERF_MAX_VIEWS][12];                 /* E_s rows 0..2 (world -> source camera)          */
  float src_ixt[ENERF_MAX_LEVELS][ENERF_MAX_VIEWS][9]; /* K_s with rows 0,1 * render_scale_l, utils.py:700-701 */
  float src_center[ENERF_MAX_VIEWS][3];               /* inv(E_s)[:3,3], utils.py:708                    */
  float tar_center[3];                                /* inv(E_t)[:3,3], utils.py:707                    */
  float near_far[2];                                  /* batch['near_far'][0]                            */
} EnerfCam;

ENERF_API int enerf_abi_version(void);
ENERF_API const char* enerf_last_error(void);

/* Replaces get_proj_mats (utils.py:35-55) and the per-view torch.inverse calls of get_img_feat
 * (utils.py:707-708).  scales [host] = n_levels x {im_feat_scale, volume_scale, render_scale}. */
ENERF_API int enerf_camera_setup(const float* src_exts /*S,4,4*/, const float* src_ixts /*S,3,3*/,
                       const float* tar_ext /*4,4*/, const float* tar_ixt /*3,3*/,
                       const float* near_far /*2*/, int n_views, int n_levels,
                       const float* scales /*[host] n_levels*3*/, EnerfCam* cam_out, void* stream);

/* Full-frame target rays generated on device: replaces the numpy
 * builder lib/datasets/enerf_utils.py:25-32,60-71 ('test' branch) and the H2D copy of rays_{i}.
 * rays (n_rows*W, 8) for image rows [row0, row0+n_rows) of the render-resolution frame of width W;
 * scale = render_scale of the level (K rows 0,1 are multiplied by it). */
ENERF_API int enerf_generate_rays(const float* tar_ext /*4,4*/, const float* tar_ixt /*3,3*/, float scale, int W, int row0,
                                  int n_rows, float* rays, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FeatureNet.forward (feature_net.py:27-36) + Network.forward_feat (network.py:58-67).
 * src_inps (S,3,H,W) NCHW in [-1,1].  H, W multiples of 4.
 * Outputs (channels-last): feat_l0 (S,H/4,W/4,32), feat_l1 (S,H/2,W/2,16), feat_l2 (S,H,W,8).
 * weights [host array, 22 device pointers], each conv as {w [tap][cin][cout] BN-folded, bias[cout]}:
 *   conv0.0, conv0.1, conv1.0, conv1.1, conv2.0, conv2.1, toplayer, lat1, lat0, smooth1, smooth0.
 * tensor_cores != 0: every layer with cin % 8 == 0 (conv0.1, conv1.0, conv1.1, conv2.0, conv2.1,
 * toplayer, smooth1, smooth0) runs as a tcgen05 implicit GEMM (TF32 operands) and takes its w in
 * the enerf_tc_conv stage layout (packing.pack_tc_conv); conv0.0 (cin 3) and the fused laterals
 * keep the [tap][cin][cout] fp32 layout and kernels.
 * part: 0 = whole net; 1 = trunk only (conv0.0..toplayer, writes feat_l0); 2 = pyramid tail only
 * (laterals + smooth convs, writes feat_l1 / feat_l2 from the trunk's workspace tensors) -- the
 * host runs part 2 on a second stream concurrently with the level-0 cost-volume chain.
 */
ENERF_API size_t enerf_feature_net_workspace_bytes(int n_views, int H, int W);
ENERF_API int enerf_feature_net(const float* const* weights, int n_weights, const float* src_inps, int n_views,
                      int H, int W, float* feat_l0, float* feat_l1, float* feat_l2, void* workspace,
                      size_t workspace_bytes, int tensor_cores, int part, void* stream);
/* The same, plus (img_feat_rgb != NULL, part 0 | 2) the (S,H,W,12) records [level-2 features | rgb * 0.5 + 0.5 | 0] that
 * enerf_pack_img_feat(feat_l2, 8, src_inps, ..., H, W) would build -- written by the fused lat0 + smooth0 launch's epilogue on the
 * tensor-core path (no extra kernel), by the pack kernel otherwise; bit-identical either way.  feat_l2 is still written. */
ENERF_API int enerf_feature_net_packed(const float* const* weights, int n_weights, const float* src_inps, int n_views,
                      int H, int W, float* feat_l0, float* feat_l1, float* feat_l2, float* img_feat_rgb, void* workspace,
                      size_t workspace_bytes, int tensor_cores, int part, void* stream);

/* cat(im_feat, unpreprocess(src_inps)) of render_rays (network.py:28-34, utils.py:605-612):
 * out (S,Hr,Wr,Cpad) with channels [0,C) = feat (must already be at Hr x Wr), [C,C+3) = rgb*0.5+0.5
 * bilinearly resized (align_corners) from (H,W) to (Hr,Wr), remaining channels 0.  Cpad = C+4. */
ENERF_API int enerf_pack_img_feat(const float* feat /*S,Hr,Wr,C*/, int C, const float* src_inps /*S,3,H,W*/,
                        int n_views, int H, int W, int Hr, int Wr, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * get_depth_values (utils.py:98-151): per-pixel plane end points of a level.
 *   level 0 (prev_depth == NULL): ends = batch near_far broadcast.
 *   level >0: bilinear (align_corners) up-sampling of the previous level's depth/std/near_far
 *             (hp,wp)->(h,w), [d+s, d-s] clamped into prev near_far and inverted (prev level is in
 *             disparity space, utils.py:122-128).
 * Writes ends (2,h,w) = metric depth of plane 0 / plane D-1 BEFORE the linspace evaluation, and
 * near_far_out (2,h,w) = what the reference returns as near_far (utils.py:148-150). */
ENERF_API int enerf_depth_hypotheses(const EnerfCam* cam, const float* prev_depth /*hp,wp or NULL*/,
                           const float* prev_std, const float* prev_near_far /*2,hp,wp*/, int hp, int wp,
                           int h, int w, int D, int depth_inv, float* ends, float* near_far_out, void* stream);

/* build_feature_volume (utils.py:322-349) = homo_warp (utils.py:57-95) over S views fused with
 * the variance (utils.py:337-345); the S warped volumes and the sampling grid are never stored.
 * feat (S,hs,ws,C) channels-last, C in {8,16,32}; variance out (D,h,w,C). */
ENERF_API int enerf_cost_volume(const EnerfCam* cam, int level, const float* feat, int n_views, int C, int hs, int ws,
                      const float* ends /*2,h,w*/, int D, int h, int w, int depth_inv, float* variance,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * MinCostRegNet.forward (cost_reg_net.py:75-86; deep=0) / CostRegNet.forward (:35-48; deep=1).
 * variance (D,h,w,in_ch) -> feat_vol (D,h,w,8) [skipped when NULL], depth_prob (D,h,w).
 * D % 4 == 0 (deep: % 8), h,w likewise.  weights [host array]: for each layer in the order
 * conv0..conv4[,conv5,conv6,conv7],conv9,conv11 : {w, bias}; then ONE head tensor (no bias):
 * w [tap][8][9] with output channel 8 = depth_conv and 0..7 = feat_conv when feat_vol != NULL, or
 * w [tap][8][1] (depth_conv only) when feat_vol == NULL.  Forward convs: w [tap][cin][cout];
 * transposed convs: w [tap][cin][cout] with tap = (kz*3+ky)*3+kx of the ConvTranspose3d kernel.
 * tensor_cores != 0: every layer runs on tcgen05 (TF32) and takes w in the enerf_tc_conv stage
 * layout (packing.pack_tc_conv; transposed layers packing.pack_tc_deconv). */
ENERF_API size_t enerf_cost_reg_workspace_bytes(int deep, int D, int h, int w);
ENERF_API int enerf_cost_reg(const float* const* weights, int n_weights, int deep, int in_ch, const float* variance,
                   int D, int h, int w, float* feat_vol, float* depth_prob, void* workspace,
                   size_t workspace_bytes, int tensor_cores, void* stream);

/* depth_regression (utils.py:658-663): softmax over D, expectation and std of the plane values
 * (in disparity when depth_inv).  depth_mvs = 1/depth when depth_inv else depth (network.py:105-108). */
ENERF_API int enerf_depth_regress(const float* depth_prob /*D,h,w*/, const float* ends /*2,h,w*/, int D, int h, int w,
                        int depth_inv, float* depth, float* std, float* depth_mvs, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The fused ray stage: build_rays (utils.py:390-420) + sample_along_depth (:422-441) +
 * get_vox_feat (:456-458) + get_img_feat (:689-722) + NeRF/Agg MLP (nerf.py:29-43,74-89) +
 * raw2outputs (utils.py:571-603), i.e. Network.render_rays (network.py:24-43) without any of the
 * (N,Ns,S,C) intermediates.  Rays are independent: a rank renders any contiguous slice.
 *   rays (n_rays,8): origin, dir (un-normalised), u, v at render resolution (Hr,Wr).
 *   depth/std (hv,wv), near_far (2,hv,wv): level outputs at volume resolution.
 *   feat_vol (D,hv,wv,8); img_feat_rgb (S,Hr,Wr,feat_ch+4) from enerf_pack_img_feat.
 *   weights [host array, 16 device pointers]: view_fc{w[4][fc],b}, global_fc{w[3fc][32],b},
 *     agg_w_fc{w[32],b}, fc{w[32][16],b}, lr0{w[24][64],b}, sigma{w[64],b},
 *     color0{w[88+fc+4][64],b}, color2{w[64],b}   (all transposed to [in][out]).
 *   outputs: rgb (n_rays,3), depth (n_rays), weights (n_rays,num_samples).
 * feat_ch (without the 3 rgb channels) in {8,32}; n_views in [2,ENERF_MAX_VIEWS];
 * num_samples in [1,8].
 *
 * ABI 2 additions:
 *   vol_row0 / vol_rows: feat_vol holds only rows [vol_row0, vol_row0 + vol_rows) of the volume, i.e. it is a
 *     (D,vol_rows,wv,8) crop (the full grid is 0 / hv).  The row-band multi-GPU layout regularises a band + halo
 *     of the cost volume per rank (enerf_b200/dist.py); the rays passed must only touch resident rows.
 *   n_rays_dev: optional DEVICE int.  When non-NULL the launch is sized for n_rays (an upper bound) and rays
 *     >= *n_rays_dev are skipped: the masked path (network_human) then needs no host read-back of the count. */
ENERF_API int enerf_render_rays(const EnerfCam* cam, int level, const float* const* weights, int n_weights,
                      const float* rays, int n_rays, const float* depth, const float* std,
                      const float* near_far, int hv, int wv, const float* feat_vol, int D, int vol_row0, int vol_rows,
                      const float* img_feat_rgb, int n_views, int Hr, int Wr, int feat_ch,
                      int num_samples, int depth_inv, int white_bkgd, int viewdir_agg, const int* n_rays_dev,
                      float* out_rgb, float* out_depth, float* out_weights, void* stream);

/* Masked-ray path (network_human.py:90-107).
 * enerf_mask_compact: order-preserving compaction of the rays whose mask element is non-zero
 *   (mask: n elements of elem_size bytes -- uint8/bool, int32, int64 ...): idx_out[i] = original ray
 *   index of the i-th kept ray, rays_out[i] = rays[idx_out[i]], *count_out = number kept (device int).
 *   == `rays[mask_at_box]` (network_human.py:92).  n <= 4096*1024.
 * enerf_scatter_rows: dst[idx[i]][0..C) = src[i][0..C) for i < m  == `rgb[mask_at_box] = ...` (:105);
 *   the caller zero-fills dst. */
ENERF_API size_t enerf_mask_compact_workspace_bytes(int n);
ENERF_API int enerf_mask_compact(const void* mask, int elem_size, const float* rays, int n, int* idx_out, float* rays_out,
                                 int* count_out, void* workspace, size_t workspace_bytes, void* stream);
/* m_dev (ABI 2): optional DEVICE row count, rows >= *m_dev are skipped (m is then the launch's upper bound) and,
 * as the reference does (`if mask_at_box.sum() > 1`, network_human.py:104), nothing is scattered unless *m_dev > 1. */
ENERF_API int enerf_scatter_rows(const float* src, const int* idx, int m, const int* m_dev, int C, float* dst, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layered ("composite") rendering, lib/networks/enerf/network_composite.py (SURVEY.md section 8f
 * row f2): every foreground layer owns a bbox window of the frame, its own near/far, cost-volume
 * regulariser and NeRF; a full-frame background layer is merged in behind them.
 * `window` / `boxes` are HOST int arrays {x, y, w, h}.
 *
 * enerf_depth_hypotheses_layer: enerf_depth_hypotheses with the layer's own [near, far] (device, 2
 *   floats) at the first level (get_depth_values_composite, utils.py:153-214).
 * enerf_cost_volume_window: the variance volume over the window only; `ends` stays (2,h,w) over the
 *   full grid, variance is (D,hc,wc,C)  (build_feature_volume_composite / homo_warp_composite,
 *   utils.py:249-320).
 * enerf_depth_regress_window: depth_regression of a (D,hc,wc) probability volume zero-padded to the
 *   (h,w) grid (network_composite.py:102-103) -> full-grid depth / std.
 * enerf_render_rays_raw: the fused ray stage of enerf_render_rays WITHOUT compositing and without the
 *   voxel feature (nerf_.NeRF, nerf_.py:29-43; pass its weights through packing.pack_nerf_novox):
 *   `rays` is the full (Hr*Wr,8) frame, the launch covers the window (build_rays_composite,
 *   utils.py:216-247), and sample k of pixel q goes to out_raw[(q*out_stride + out_off + k)*4 .. +4)
 *   = (r,g,b,sigma), out_z[q*out_stride + out_off + k] = metric z (network_composite.py:48-51).
 * enerf_composite_layers: parse_layer + raw2outputs_composite (utils.py:875-942).  raw/z as written by
 *   enerf_render_rays_raw with out_stride = n_fg_layers*ns_fg + ns_bg (layer l at offset l*ns_fg, the
 *   background last); samples outside their layer's box count as zeros.  Outputs: rgb (HW,3),
 *   depth (HW), weights (HW,n_total), net_output (HW,n_total,4) in composited order, z_vals
 *   (HW,n_fg) before sorting, idx (HW,n_fg) int64 sort permutation (written when n_fg_layers > 1,
 *   else may be NULL).  n_fg_layers in [1,8], n_fg_layers*ns_fg <= 32. */
ENERF_API int enerf_depth_hypotheses_layer(const float* layer_near_far, const float* prev_depth, const float* prev_std,
                                           const float* prev_near_far, int hp, int wp, int h, int w, int D, int depth_inv,
                                           float* ends, float* near_far_out, void* stream);
ENERF_API int enerf_cost_volume_window(const EnerfCam* cam, int level, const float* feat, int n_views, int C, int hs, int ws,
                                       const float* ends, int D, int h, int w, const int* window, int depth_inv, float* variance,
                                       void* stream);
ENERF_API int enerf_depth_regress_window(const float* depth_prob, const int* window, const float* ends, int D, int h, int w,
                                         int depth_inv, float* depth, float* std, float* depth_mvs, void* stream);
ENERF_API int enerf_render_rays_raw(const EnerfCam* cam, int level, const float* const* weights, int n_weights, const float* rays,
                                    const int* window, const float* depth, const float* std, const float* near_far, int hv, int wv,
                                    const float* img_feat_rgb, int n_views, int Hr, int Wr, int feat_ch, int num_samples,
                                    int depth_inv, int viewdir_agg, float* out_raw, float* out_z, int out_stride, int out_off,
                                    void* stream);
/* enerf_render_rays_raw with the MLP on the tensor cores (wblob: packing.pack_nerf_tc_novox); same
 * support matrix as enerf_render_rays_tc (feat_ch 8, n_views 2..8, num_samples in {1,2,4,8}). */
ENERF_API int enerf_render_rays_raw_tc(const EnerfCam* cam, int level, const float* wblob, const float* rays, const int* window,
                                       const float* depth, const float* std, const float* near_far, int hv, int wv,
                                       const float* img_feat_rgb, int n_views, int Hr, int Wr, int feat_ch, int num_samples,
                                       int depth_inv, int viewdir_agg, float* out_raw, float* out_z, int out_stride, int out_off,
                                       void* stream);
ENERF_API int enerf_composite_layers(const float* raw, const float* z, int Hr, int Wr, int n_fg_layers, int ns_fg, int ns_bg,
                                     const int* boxes, float* rgb, float* depth, float* weights, float* net_output,
                                     long long* idx, float* z_vals, void* stream);

/* Device-side consumers of the rendered frame.
 * enerf_psnr_accumulate: acc[0] += sum (pred-gt)^2 over the pixels whose mask element is non-zero
 *   (mask == NULL: all pixels) x 3 channels, acc[1] += number of values; PSNR = 10 log10(acc[1]/acc[0])
 *   (lib/evaluators/enerf.py:45-71, skimage psnr with data_range 1).  acc: 2 doubles, caller-zeroed.
 * enerf_pack_rgb8: (H*W,3) float rgb -> (H,W,3) uint8 = trunc(clamp(x,0,1)*255), optional vertical
 *   flip (gui_human.py:88-91). */
ENERF_API int enerf_psnr_accumulate(const float* pred, const float* gt, const void* mask, int mask_elem_size,
                                    long long n_pixels, double* acc, void* stream);
ENERF_API int enerf_pack_rgb8(const float* rgb, int H, int W, int flip_vertical, unsigned char* out, void* stream);

/* Tensor-core variant of enerf_render_rays (same stage, same arguments, same outputs): the MLP
 * contractions run as tcgen05.mma kind::tf32 with accumulators in TMEM, 128 sample points per CTA.
 * feat_ch == 8, n_views in [2,8], num_samples in {1,2,4,8}.  wblob = one packed device buffer of
 * 10,392 floats (enerf_b200/packing.py::pack_nerf_tc: TF32-rounded B operands in the K-major
 * 16-byte-chunk layout of csrc/tc.cuh, followed by the fp32 bias / 1-wide vectors). */
ENERF_API int enerf_render_rays_tc(const EnerfCam* cam, int level, const float* wblob, const float* rays, int n_rays,
                                   const float* depth, const float* std, const float* near_far, int hv, int wv,
                                   const float* feat_vol, int D, int vol_row0, int vol_rows, const float* img_feat_rgb,
                                   int n_views, int Hr, int Wr, int feat_ch, int num_samples, int depth_inv, int white_bkgd,
                                   int viewdir_agg, const int* n_rays_dev, float* out_rgb, float* out_depth,
                                   float* out_weights, void* stream);

/* One convolution layer on the tensor cores (tcgen05 implicit GEMM, csrc/tc_conv.cuh): the building
 * block enerf_feature_net / enerf_cost_reg use in TF32 mode, exported for layer-level parity tests.
 *   kind 0: convolution KD x KH x KH (KD in {1,3}, KH in {1,3}; KH = 5 with KD = 1), zero padding K/2,
 *           stride 1 or 2 (stride 2 strides H and W, and D when KD = 3; (D,H,W) = INPUT extent, even).
 *   kind 1: ConvTranspose3d(k3, s2, p1, output_padding 1) in sub-pixel form; (D,H,W) is the INPUT grid.
 *   mode 0: out[pix][out_coff..+cout) = act(acc + bias);  1: head (8 feat -> out, 1 prob -> out2);
 *        2: transposed conv, out = skip + acc + bias;     3: single channel -> out.
 * in (D,H,W,cin) channels-last, cin % 8 == 0.  wpack: enerf_b200/packing.py::pack_tc_conv /
 * pack_tc_deconv ([cin/8][tap][2][N][4], TF32-rounded, N = cout (x8 for kind 1) padded to 16). */
ENERF_API int enerf_tc_conv(int kind, int KD, int KH, int stride, int cin, int cout, int mode, int relu, const float* in, int D, int H,
                            int W, const float* wpack, const float* bias, const float* skip, float* out, float* out2,
                            int out_cstride, int out_coff, void* stream);

/* Diagnostic: which kernel serves enerf_render_rays_tc / enerf_render_rays_raw_tc.  0 = auto (currently the single-role
 * kernel, csrc/render_rays_tc.cuh, 2-8 views), 1 = the single-role kernel, 2 = the warp-specialised kernel
 * (csrc/render_rays_ws.cu: two gather warpgroups + a consumer warpgroup, factored MLP) for 2-3 source views. */
ENERF_API int enerf_render_rays_tc_select(int impl);

/* Diagnostic / tuning of the persistent TMA-fed convolution kernel (csrc/tc_conv2.cu) behind enerf_tc_conv and the
 * conv stacks: impl 0 = auto (it takes every layer whose weights fit in shared memory),
 * 1 = csrc/tc_conv.cu's kernel only, 2 = same as 0, 3 = auto without the stride-2 layers (their phase tiles are TMA boxes
 * with element stride 2);
 * nmma = MMA-issuing warps per CTA (1|2, 0 = default 2); ctas_per_sm (1|2, 0 = default 2);
 * tz, ty, kbc (8|16|32), slots: forced tile / K-block width / ring depth, 0 = built-in choice. */
ENERF_API int enerf_tc_conv2_tune(int impl, int nmma, int ctas_per_sm, int tz, int ty, int kbc, int slots);
/* Diagnostic: in-kernel timeline of csrc/tc_conv2.cu (CTA 0, first 16 tiles): buf = 3 x 16 x 8 u64 of %globaltimer ns
 * (roles producer / MMA warp 0 / epilogue row 0; layout at the definition); NULL switches it off. */
ENERF_API int enerf_tc_conv2_debug(unsigned long long* buf);
/* ... the same, stamped only by the fused-lateral launch of enerf_feature_net (role 0 = computing producer thread 0). */
ENERF_API int enerf_tc_conv2_debug_lateral(unsigned long long* buf);
/* The launch geometry csrc/tc_conv2.cu would use for a layer ((D,H,W) = its row grid) on a device with n_sm SMs, computed
 * without touching a GPU, so the CPU test suite can emulate the kernel from it (tests/test_host_cpu.py).  out: 67 ints,
 * layout documented at the definition.  Returns ENERF_EUNSUPPORTED when the layer stays on csrc/tc_conv.cu. */
ENERF_API int enerf_tc_conv2_plan(int kind, int KD, int KH, int stride, int cin, int cout, int mode, int D, int H, int W, int fold, int lateral,
                                  int n_sm, int* out);
/* Diagnostic: 1 (default) = enerf_feature_net computes lat0 (1x1 lateral + bilinear x2 + add, feature_net.py:31-35) inside
 * smooth0's producer warps on the tensor-core path (source tiles staged by TMA; the 126 MB 32-channel map is never written);
 * 0 = separate lateral kernel + plain smooth0 (bit-identical features); 4 | 6 | 8 = fused with that many computing producer
 * warps per CTA (1 = the default, 6). */
ENERF_API int enerf_tc_conv2_fuse_lateral(int on);

/* Diagnostic: when buf != NULL, CTA (0,0,0) of every later enerf_tc_conv-family launch writes 64
 * %globaltimer phase stamps (ns) into buf (device memory, 64 x u64).  NULL switches it off. */
ENERF_API int enerf_tc_conv_debug(unsigned long long* buf);

/* Diagnostic / tuning: force the tile (tz x ty rows of 32 positions) and the kx-tap folding (0 off, 1 on
 * where applicable, -1 built-in) of every later enerf_tc_conv-family launch; (0, 0, -1) restores the
 * built-in choice.  Weights must be packed for the folding in force (packing.pack_tc_conv fold_kx). */
ENERF_API int enerf_tc_conv_tune(int tz, int ty, int fold);

/* Which convolutions carry their three kx taps in the MMA's N dimension (3x fewer tcgen05.mma, a row-shift exchange in the
 * epilogue): 0 = stride-1 3x3x3 layers with 8 output channels + the depth-only head, 1 = + the feat/prob head, 2 (default) =
 * + stride-1 3x3 2-D layers with 8 output channels.  The packed weights must follow the same rule
 * (enerf_b200/packing.py::FOLD_RULE; enerf_b200.capi.tc_conv_fold_rule sets both and the caller re-packs). */
ENERF_API int enerf_tc_conv_fold_rule(int level);

/* Test hook (no GPU): n / d (d >= 1, n < 2^31) through the multiply-high constants the kernels use for divisions by run-time extents. */
ENERF_API unsigned enerf_fastdiv_check(unsigned d, unsigned n);

/* Diagnostic: like enerf_tc_conv_debug for enerf_render_rays_tc (32 x u64: 16 stamps for each of
 * the first two tiles of CTA 0). */
ENERF_API int enerf_render_rays_debug(unsigned long long* buf);

/* Diagnostic microbenchmark: ns for n_mma back-to-back tcgen05.mma (M=128, K=8 tf32) with the given
 * operand layout (0 none, 2/4/6 = 128/64/32-byte swizzle), N, and `accs` accumulators cycled. */
ENERF_API int enerf_tc_mma_bench(int layout, int N, int n_mma, int accs, unsigned long long* out_ns, void* stream);

/* Diagnostic: D[128 x N] = A[128 x K] * B[N x K]^T on the tensor cores (tcgen05.mma kind::tf32,
 * accumulator in TMEM) through the same descriptor helpers the fused kernels use (csrc/tc.cuh).
 * K multiple of 8 (<=128), N multiple of 16 (<=256); A, B row-major.  No reference counterpart. */
ENERF_API int enerf_tc_selftest(const float* A, const float* B, int K, int N, float* D, void* stream);

/* Diagnostic: the convention enerf_tc_conv's TMA-fed kernel rests on.  A [rows x Kf] (Kf = 8|16|32 floats =
 * a 32|64|128-byte swizzle span) is loaded by ONE TMA box with the matching swizzle; D[128 x N] =
 * A[row_off : row_off+128] * B[N x Kf]^T is computed with a K-major SWIZZLED operand descriptor whose start
 * address is advanced by row_off whole rows.  bo_mode: the descriptor's base_offset field, 0 = zero,
 * 1 = (start >> 7) & 7.  No reference counterpart. */
ENERF_API int enerf_tc_swz_selftest(const float* A, int rows, int Kf, const float* B, int N, int row_off, int bo_mode, float* D, void* stream);

/* Diagnostic: TMA box rate.  `grid` CTAs each stream `iters` halo boxes {C, tx+2, ty+2, tz+2 (1 if tz == 1)}
 * of the channels-last fp32 tensor x (D,H,W,C), C = 8|16|32, with `depth` boxes in flight. */
/* Diagnostic: MMA issue rate with several issuing warps per CTA (1..4) and several CTAs per SM (grid, pad_bytes of extra
 * dynamic shared memory set the co-residency); out_ns[cta*4 + warp] = that issuer's elapsed ns. */
ENERF_API int enerf_tc_mma_bench2(int layout, int N, int n_mma, int n_issuers, int ksteps, int grid, int pad_bytes, unsigned long long* out_ns,
                                  void* stream);
/* Diagnostic: cost of a TMEM read (tcgen05.ld.32x32b.x{8,16,32} + wait::ld, n_ld iterations by warps 4-7) while the tensor
 * pipe is idle (mode 0) or fed with n_mma back-to-back MMAs by one (1) or two (2) issuing warps.  out_ns[cta*4 + warp]. */
ENERF_API int enerf_tc_ldtm_bench(int mode, int N, int n_mma, int n_ld, int cols, int grid, unsigned long long* out_ns, void* stream);
ENERF_API int enerf_tma_box_bench(const float* x, int D, int H, int W, int C, int tx, int ty, int tz, int depth, int iters, int grid,
                                  float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ENERF_B200_H */
