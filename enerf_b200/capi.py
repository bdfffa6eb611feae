"""ctypes binding of libenerf_b200.so (include/enerf_b200.h).  Thin on purpose: unpack torch tensors
to raw device pointers, pass the current CUDA stream, turn error codes into Python exceptions.

There is no fallback of any kind: if the library is missing or cannot be loaded the import of the
compute path raises (ImportError), and every entry point raises on a non-CUDA tensor.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libenerf_b200.so")

MAX_VIEWS, MAX_LEVELS = 8, 4
CAM_FLOATS = MAX_LEVELS * MAX_VIEWS * 12 + MAX_VIEWS * 12 + MAX_LEVELS * MAX_VIEWS * 9 + MAX_VIEWS * 3 + 3 + 2

ERRORS = {-1: ValueError, -2: RuntimeError, -3: RuntimeError, -4: NotImplementedError}

_lib = None

# kernels launched by this process through the library (each wrapper adds what its entry point
# launches; bench.py reports the per-step count as `gpu_launches`)
LAUNCHES = 0

_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
_SIGS = {
    "enerf_abi_version": (_i, []),
    "enerf_last_error": (ctypes.c_char_p, []),
    "enerf_camera_setup": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, ctypes.POINTER(ctypes.c_float), _vp, _vp]),
    "enerf_generate_rays": (_i, [_vp, _vp, ctypes.c_float, _i, _i, _i, _vp, _vp]),
    "enerf_feature_net_workspace_bytes": (_sz, [_i, _i, _i]),
    "enerf_feature_net": (_i, [ctypes.POINTER(_vp), _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    "enerf_feature_net_packed": (_i, [ctypes.POINTER(_vp), _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    "enerf_pack_img_feat": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "enerf_depth_hypotheses": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "enerf_cost_volume": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "enerf_cost_reg_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "enerf_cost_reg": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "enerf_depth_regress": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "enerf_render_rays": (_i, [_vp, _i, ctypes.POINTER(_vp), _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _i, _i,
                               _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "enerf_depth_hypotheses_layer": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "enerf_cost_volume_window": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, ctypes.POINTER(_i), _i, _vp, _vp]),
    "enerf_depth_regress_window": (_i, [_vp, ctypes.POINTER(_i), _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "enerf_render_rays_raw": (_i, [_vp, _i, ctypes.POINTER(_vp), _i, _vp, ctypes.POINTER(_i), _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i,
                                   _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "enerf_render_rays_raw_tc": (_i, [_vp, _i, _vp, _vp, ctypes.POINTER(_i), _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i,
                                      _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "enerf_composite_layers": (_i, [_vp, _vp, _i, _i, _i, _i, _i, ctypes.POINTER(_i), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "enerf_mask_compact_workspace_bytes": (_sz, [_i]),
    "enerf_mask_compact": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "enerf_scatter_rows": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp]),
    "enerf_psnr_accumulate": (_i, [_vp, _vp, _vp, _i, ctypes.c_longlong, _vp, _vp]),
    "enerf_pack_rgb8": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "enerf_render_rays_tc": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _i, _i,
                                  _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "enerf_render_rays_tc_select": (_i, [_i]),
    "enerf_tc_conv": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "enerf_tc_conv_debug": (_i, [_vp]),
    "enerf_tc_conv_tune": (_i, [_i, _i, _i]),
    "enerf_tc_conv_fold_rule": (_i, [_i]),
    "enerf_fastdiv_check": (ctypes.c_uint, [ctypes.c_uint, ctypes.c_uint]),
    "enerf_tc_conv2_tune": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "enerf_tc_conv2_fuse_lateral": (_i, [_i]),
    "enerf_tc_conv2_debug": (_i, [_vp]),
    "enerf_tc_conv2_debug_lateral": (_i, [_vp]),
    "enerf_tc_conv2_plan": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_i)]),
    "enerf_tc_mma_bench": (_i, [_i, _i, _i, _i, _vp, _vp]),
    "enerf_render_rays_debug": (_i, [_vp]),
    "enerf_tc_selftest": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "enerf_tc_swz_selftest": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "enerf_tc_mma_bench2": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "enerf_tc_ldtm_bench": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp]),
    "enerf_tma_box_bench": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}
EXPORTS = tuple(_SIGS)


def lib():
    """Load (once) and return the ctypes handle; raises ImportError when the CUDA library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -m enerf_b200.build` "
                              "(there is no CPU fallback for the render path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.enerf_abi_version() != 2:
            raise ImportError("libenerf_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def _check(rc, what, launches=1):
    global LAUNCHES
    LAUNCHES += launches
    if rc != 0:
        msg = lib().enerf_last_error().decode("utf-8", "replace")
        raise ERRORS.get(rc, RuntimeError)(f"{what} failed ({rc}): {msg}")


def ptr(t, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError("null tensor")
    if not t.is_cuda:
        raise ValueError("enerf_b200 kernels take CUDA tensors only (no CPU fallback)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"expected contiguous float32, got {t.dtype} contiguous={t.is_contiguous()}")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr_array(tensors):
    arr = (_vp * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = ptr(t, allow_none=True)
    return arr


# ---- one wrapper per entry point -------------------------------------------------------------
def camera_setup(src_exts, src_ixts, tar_ext, tar_ixt, near_far, scales, cam):
    S = src_exts.shape[0]
    n_levels = len(scales)
    flat = (ctypes.c_float * (n_levels * 3))(*[float(x) for lv in scales for x in lv])
    _check(lib().enerf_camera_setup(ptr(src_exts), ptr(src_ixts), ptr(tar_ext), ptr(tar_ixt), ptr(near_far), S, n_levels,
                                    flat, ptr(cam), stream()), "enerf_camera_setup")


def feature_net_workspace_bytes(S, H, W):
    return lib().enerf_feature_net_workspace_bytes(S, H, W)


def feature_net(weights, src_inps, feat_l0, feat_l1, feat_l2, workspace, tensor_cores=False, part=0, img_feat_rgb=None):
    """img_feat_rgb: optional (S,H,W,12) tensor that receives the [level-2 features | rgb | 0] records of pack_img_feat, written by
    the fused lat0 + smooth0 launch itself on the tensor-core path (by the pack kernel otherwise); filled by part 0 and part 2."""
    S, _, H, W = src_inps.shape
    fused = bool(tensor_cores and part != 1 and _FUSE_LAT and _CONV_IMPL != 1)
    launches = {0: 11, 1: 7, 2: 4}[part] - (1 if fused else 0) + (1 if (img_feat_rgb is not None and part != 1 and not fused) else 0)
    if img_feat_rgb is not None and tuple(img_feat_rgb.shape) != (S, H, W, 12):
        raise ValueError(f"img_feat_rgb must be ({S},{H},{W},12), got {tuple(img_feat_rgb.shape)}")
    _check(lib().enerf_feature_net_packed(ptr_array(weights), len(weights), ptr(src_inps), S, H, W, ptr(feat_l0), ptr(feat_l1),
                                          ptr(feat_l2), ptr(img_feat_rgb, True), workspace.data_ptr(),
                                          workspace.numel() * workspace.element_size(), int(tensor_cores), part, stream()),
           "enerf_feature_net_packed", launches=launches)


def pack_img_feat(feat, src_inps, out):
    S, Hr, Wr, C = feat.shape
    H, W = src_inps.shape[-2:]
    _check(lib().enerf_pack_img_feat(ptr(feat), C, ptr(src_inps), S, H, W, Hr, Wr, ptr(out), stream()), "enerf_pack_img_feat")


def depth_hypotheses(cam, prev_depth, prev_std, prev_nf, h, w, D, depth_inv, ends, nf_out):
    hp, wp = (prev_depth.shape[-2:] if prev_depth is not None else (0, 0))
    _check(lib().enerf_depth_hypotheses(ptr(cam), ptr(prev_depth, True), ptr(prev_std, True), ptr(prev_nf, True), hp, wp, h, w, D,
                                        int(depth_inv), ptr(ends), ptr(nf_out), stream()), "enerf_depth_hypotheses")


def cost_volume(cam, level, feat, ends, D, h, w, depth_inv, variance):
    S, hs, ws, C = feat.shape
    _check(lib().enerf_cost_volume(ptr(cam), level, ptr(feat), S, C, hs, ws, ptr(ends), D, h, w, int(depth_inv), ptr(variance),
                                   stream()), "enerf_cost_volume")


def cost_reg_workspace_bytes(deep, D, h, w):
    return lib().enerf_cost_reg_workspace_bytes(int(deep), D, h, w)


def cost_reg(weights, deep, variance, feat_vol, depth_prob, workspace, tensor_cores=False):
    D, h, w, C = variance.shape
    _check(lib().enerf_cost_reg(ptr_array(weights), len(weights), int(deep), C, ptr(variance), D, h, w, ptr(feat_vol, True),
                                ptr(depth_prob), workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                                int(tensor_cores), stream()),
           "enerf_cost_reg", launches=(11 if deep else 8))


def depth_regress(depth_prob, ends, depth_inv, depth, std, depth_mvs):
    D, h, w = depth_prob.shape
    _check(lib().enerf_depth_regress(ptr(depth_prob), ptr(ends), D, h, w, int(depth_inv), ptr(depth), ptr(std), ptr(depth_mvs, True),
                                     stream()), "enerf_depth_regress")


def _count_ptr(t):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.int32:
        raise ValueError("device-side counts must be CUDA int32 tensors")
    return t.data_ptr()


def render_rays(cam, level, weights, rays, depth, std, near_far, feat_vol, img_feat_rgb, feat_ch, num_samples, depth_inv,
                white_bkgd, viewdir_agg, out_rgb, out_depth, out_weights, vol_row0=0, n_rays_dev=None):
    """feat_vol (D, rows, wv, 8): rows [vol_row0, vol_row0 + rows) of the level's volume (the full grid by default)."""
    n_rays = rays.shape[0]
    hv, wv = depth.shape[-2:]
    D, vol_rows = feat_vol.shape[0], feat_vol.shape[1]
    S, Hr, Wr, _ = img_feat_rgb.shape
    _check(lib().enerf_render_rays(ptr(cam), level, ptr_array(weights), len(weights), ptr(rays), n_rays, ptr(depth), ptr(std),
                                   ptr(near_far), hv, wv, ptr(feat_vol), D, vol_row0, vol_rows, ptr(img_feat_rgb), S, Hr, Wr, feat_ch,
                                   num_samples, int(depth_inv), int(white_bkgd), int(viewdir_agg), _count_ptr(n_rays_dev), ptr(out_rgb),
                                   ptr(out_depth), ptr(out_weights), stream()), "enerf_render_rays")


def _ints(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def depth_hypotheses_layer(layer_near_far, prev_depth, prev_std, prev_nf, h, w, D, depth_inv, ends, nf_out):
    hp, wp = (prev_depth.shape[-2:] if prev_depth is not None else (0, 0))
    _check(lib().enerf_depth_hypotheses_layer(ptr(layer_near_far, True), ptr(prev_depth, True), ptr(prev_std, True), ptr(prev_nf, True),
                                              hp, wp, h, w, D, int(depth_inv), ptr(ends), ptr(nf_out), stream()),
           "enerf_depth_hypotheses_layer")


def cost_volume_window(cam, level, feat, ends, D, h, w, window, depth_inv, variance):
    S, hs, ws, C = feat.shape
    _check(lib().enerf_cost_volume_window(ptr(cam), level, ptr(feat), S, C, hs, ws, ptr(ends), D, h, w, _ints(window), int(depth_inv),
                                          ptr(variance), stream()), "enerf_cost_volume_window")


def depth_regress_window(depth_prob, window, ends, depth_inv, depth, std, depth_mvs=None):
    D = depth_prob.shape[0]
    h, w = ends.shape[-2:]
    _check(lib().enerf_depth_regress_window(ptr(depth_prob), _ints(window), ptr(ends), D, h, w, int(depth_inv), ptr(depth), ptr(std),
                                            ptr(depth_mvs, True), stream()), "enerf_depth_regress_window")


def render_rays_raw(cam, level, weights, rays, window, depth, std, near_far, img_feat_rgb, feat_ch, num_samples, depth_inv,
                    viewdir_agg, out_raw, out_z, out_off):
    hv, wv = depth.shape[-2:]
    S, Hr, Wr, _ = img_feat_rgb.shape
    if rays.shape[0] != Hr * Wr:
        raise ValueError(f"render_rays_raw wants the full {Hr}x{Wr} frame of rays, got {rays.shape[0]}")
    _check(lib().enerf_render_rays_raw(ptr(cam), level, ptr_array(weights), len(weights), ptr(rays), _ints(window), ptr(depth), ptr(std),
                                       ptr(near_far), hv, wv, ptr(img_feat_rgb), S, Hr, Wr, feat_ch, num_samples, int(depth_inv),
                                       int(viewdir_agg), ptr(out_raw), ptr(out_z), out_z.shape[-1], out_off, stream()),
           "enerf_render_rays_raw", launches=(1 if window[2] * window[3] > 0 else 0))


def render_rays_raw_tc(cam, level, wblob, rays, window, depth, std, near_far, img_feat_rgb, feat_ch, num_samples, depth_inv,
                       viewdir_agg, out_raw, out_z, out_off):
    hv, wv = depth.shape[-2:]
    S, Hr, Wr, _ = img_feat_rgb.shape
    if rays.shape[0] != Hr * Wr:
        raise ValueError(f"render_rays_raw_tc wants the full {Hr}x{Wr} frame of rays, got {rays.shape[0]}")
    _check(lib().enerf_render_rays_raw_tc(ptr(cam), level, ptr(wblob), ptr(rays), _ints(window), ptr(depth), ptr(std), ptr(near_far), hv, wv,
                                          ptr(img_feat_rgb), S, Hr, Wr, feat_ch, num_samples, int(depth_inv), int(viewdir_agg),
                                          ptr(out_raw), ptr(out_z), out_z.shape[-1], out_off, stream()),
           "enerf_render_rays_raw_tc", launches=(1 if window[2] * window[3] > 0 else 0))


def composite_layers(raw, z, Hr, Wr, n_fg_layers, ns_fg, ns_bg, boxes, rgb, depth, weights, net_output, idx, z_vals):
    flat = [v for b in boxes for v in b]
    if idx is not None and (idx.dtype != torch.int64 or not idx.is_cuda or not idx.is_contiguous()):
        raise ValueError("idx must be a contiguous CUDA int64 tensor")
    _check(lib().enerf_composite_layers(ptr(raw), ptr(z), Hr, Wr, n_fg_layers, ns_fg, ns_bg, _ints(flat), ptr(rgb), ptr(depth),
                                        ptr(weights), ptr(net_output), idx.data_ptr() if idx is not None else None, ptr(z_vals), stream()),
           "enerf_composite_layers")


def tc_conv_tune(tz=0, ty=0, fold=-1):
    _check(lib().enerf_tc_conv_tune(int(tz), int(ty), int(fold)), "enerf_tc_conv_tune", launches=0)


def tc_conv_fold_rule(level=2):
    """Select which layers fold their kx taps into N, in the library AND in packing.py (weights packed before the call must be
    re-packed: Network.invalidate_packed())."""
    from . import packing
    _check(lib().enerf_tc_conv_fold_rule(int(level)), "enerf_tc_conv_fold_rule", launches=0)
    packing.FOLD_RULE = int(level)


_FUSE_LAT, _CONV_IMPL = True, 0     # mirrors of the library's switches, for the launch count only


def tc_conv2_tune(impl=0, nmma=0, ctas_per_sm=0, tz=0, ty=0, kbc=0, slots=0):
    global _CONV_IMPL
    _CONV_IMPL = int(impl)
    _check(lib().enerf_tc_conv2_tune(int(impl), int(nmma), int(ctas_per_sm), int(tz), int(ty), int(kbc), int(slots)), "enerf_tc_conv2_tune", launches=0)


PLAN_FIELDS = ("TZ TY TX IZ IY IX oz oy ox nx ny nz n_tiles sz sy sx n_phases phase_bytes kbc n_kb n_slots slot_bytes box_bytes "
               "N n_mt n_acc n_taps fold tmem_cols w_bytes xch_bytes").split()


def tc_conv2_plan(kind, KD, KH, stride, cin, cout, mode, D, H, W, fold, lateral=False, n_sm=148):
    """Launch geometry of csrc/tc_conv2.cu for a layer (no GPU needed); dict of PLAN_FIELDS + 'tap_off' (16-byte units)."""
    out = (ctypes.c_int * 67)()
    _check(lib().enerf_tc_conv2_plan(kind, KD, KH, stride, cin, cout, mode, D, H, W, int(fold), int(lateral), n_sm, out), "enerf_tc_conv2_plan", launches=0)
    plan = {k: out[i] for i, k in enumerate(PLAN_FIELDS)}
    plan["tap_off"] = [out[40 + i] for i in range(plan["n_taps"])]
    return plan


def tc_conv2_debug(buf):
    """buf: int64 CUDA tensor of 3*16*8 elements (or None to switch the stamps off)."""
    _check(lib().enerf_tc_conv2_debug(buf.data_ptr() if buf is not None else None), "enerf_tc_conv2_debug", launches=0)


def tc_conv2_debug_lateral(buf):
    """buf: int64 CUDA tensor of 3*16*8 elements (or None to switch the stamps off)."""
    _check(lib().enerf_tc_conv2_debug_lateral(buf.data_ptr() if buf is not None else None), "enerf_tc_conv2_debug_lateral", launches=0)


def tc_conv2_fuse_lateral(on=True):
    """on: False | True | 4 | 6 | 8 (fused, with that many computing producer warps; True = the default, 6)."""
    global _FUSE_LAT
    _FUSE_LAT = bool(on)
    _check(lib().enerf_tc_conv2_fuse_lateral(int(on)), "enerf_tc_conv2_fuse_lateral", launches=0)


def tc_selftest(A, B, D):
    """D[128,N] = A[128,K] @ B[N,K]^T on tcgen05 (TF32 operands, fp32 accumulate in TMEM)."""
    _check(lib().enerf_tc_selftest(ptr(A), ptr(B), A.shape[1], B.shape[0], ptr(D), stream()), "enerf_tc_selftest")


def tc_swz_selftest(A, B, D, row_off, bo_mode):
    """D[128,N] = A[row_off:row_off+128] @ B^T with A staged by a swizzled TMA box (see enerf_b200.h)."""
    _check(lib().enerf_tc_swz_selftest(ptr(A), A.shape[0], A.shape[1], ptr(B), B.shape[0], row_off, bo_mode, ptr(D), stream()), "enerf_tc_swz_selftest")


def tc_mma_bench2(layout, N, n_mma, n_issuers, ksteps=1, grid=1, pad_bytes=0):
    """Returns the (grid, 4) int64 tensor of per-issuer elapsed ns."""
    out = torch.zeros(grid * 4, dtype=torch.int64, device="cuda")
    _check(lib().enerf_tc_mma_bench2(layout, N, n_mma, n_issuers, ksteps, grid, pad_bytes, out.data_ptr(), stream()), "enerf_tc_mma_bench2")
    torch.cuda.synchronize()
    return out.view(grid, 4).cpu()


def tc_ldtm_bench(mode, N, n_mma, n_ld, cols, grid=148):
    out = torch.zeros(grid * 4, dtype=torch.int64, device="cuda")
    _check(lib().enerf_tc_ldtm_bench(mode, N, n_mma, n_ld, cols, grid, out.data_ptr(), stream()), "enerf_tc_ldtm_bench")
    torch.cuda.synchronize()
    return out.view(grid, 4).cpu()


def tma_box_bench(x, tx, ty, tz, depth, iters, grid, sink):
    Dd, H, W, C = x.shape
    _check(lib().enerf_tma_box_bench(ptr(x), Dd, H, W, C, tx, ty, tz, depth, iters, grid, ptr(sink), stream()), "enerf_tma_box_bench")


def render_rays_tc(cam, level, wblob, rays, depth, std, near_far, feat_vol, img_feat_rgb, feat_ch, num_samples, depth_inv,
                   white_bkgd, viewdir_agg, out_rgb, out_depth, out_weights, vol_row0=0, n_rays_dev=None):
    n_rays = rays.shape[0]
    hv, wv = depth.shape[-2:]
    D, vol_rows = feat_vol.shape[0], feat_vol.shape[1]
    S, Hr, Wr, _ = img_feat_rgb.shape
    _check(lib().enerf_render_rays_tc(ptr(cam), level, ptr(wblob), ptr(rays), n_rays, ptr(depth), ptr(std), ptr(near_far), hv, wv,
                                      ptr(feat_vol), D, vol_row0, vol_rows, ptr(img_feat_rgb), S, Hr, Wr, feat_ch, num_samples,
                                      int(depth_inv), int(white_bkgd), int(viewdir_agg), _count_ptr(n_rays_dev), ptr(out_rgb),
                                      ptr(out_depth), ptr(out_weights), stream()),
           "enerf_render_rays_tc")


def render_rays_tc_select(impl=0):
    _check(lib().enerf_render_rays_tc_select(int(impl)), "enerf_render_rays_tc_select", launches=0)


def tc_ray_kernel_supports(feat_ch, n_views, num_samples):
    return feat_ch == 8 and 2 <= n_views <= MAX_VIEWS and num_samples in (1, 2, 4, 8)


def tc_conv(kind, KD, KH, cout, mode, relu, x, wpack, bias, skip, out, out2=None, out_cstride=None, out_coff=0, stride=1):
    """x (D,H,W,cin) channels-last.  See enerf_tc_conv in include/enerf_b200.h."""
    D, H, W, cin = x.shape
    _check(lib().enerf_tc_conv(kind, KD, KH, stride, cin, cout, mode, int(relu), ptr(x), D, H, W, ptr(wpack), ptr(bias, True), ptr(skip, True),
                               ptr(out), ptr(out2, True), out_cstride if out_cstride is not None else cout, out_coff, stream()),
           "enerf_tc_conv")


def tc_conv_debug(buf):
    """buf: int64 CUDA tensor of 64 elements (or None to switch the phase stamps off)."""
    _check(lib().enerf_tc_conv_debug(buf.data_ptr() if buf is not None else None), "enerf_tc_conv_debug", launches=0)


def tc_mma_bench(layout, N, n_mma, accs=1):
    out = torch.zeros(1, dtype=torch.int64, device="cuda")
    _check(lib().enerf_tc_mma_bench(layout, N, n_mma, accs, out.data_ptr(), stream()), "enerf_tc_mma_bench")
    torch.cuda.synchronize()
    return int(out.item())


def render_rays_debug(buf):
    _check(lib().enerf_render_rays_debug(buf.data_ptr() if buf is not None else None), "enerf_render_rays_debug", launches=0)


def generate_rays(tar_ext, tar_ixt, scale, W, row0, n_rows, rays):
    _check(lib().enerf_generate_rays(ptr(tar_ext), ptr(tar_ixt), float(scale), W, row0, n_rows, ptr(rays), stream()), "enerf_generate_rays")


def mask_compact(mask, rays, idx_out, rays_out, count_out, workspace):
    """mask: contiguous CUDA tensor with one element per ray (any integer/bool dtype)."""
    if not mask.is_cuda or not mask.is_contiguous():
        raise ValueError("mask must be a contiguous CUDA tensor")
    n = rays.shape[0]
    _check(lib().enerf_mask_compact(mask.data_ptr(), mask.element_size(), ptr(rays), n, idx_out.data_ptr(), ptr(rays_out),
                                    count_out.data_ptr(), workspace.data_ptr(), workspace.numel() * workspace.element_size(), stream()),
           "enerf_mask_compact", launches=3)


def mask_compact_workspace_bytes(n):
    return lib().enerf_mask_compact_workspace_bytes(n)


def scatter_rows(src, idx, m, dst, m_dev=None):
    C = dst.shape[-1]
    _check(lib().enerf_scatter_rows(ptr(src) if m else None, idx.data_ptr() if m else None, m, _count_ptr(m_dev), C, ptr(dst), stream()),
           "enerf_scatter_rows")


def psnr_accumulate(pred, gt, mask, acc):
    """pred, gt (N,3) float CUDA; mask (N,) any integer/bool dtype or None; acc: 2 float64 (zeroed by the caller)."""
    n = pred.numel() // 3
    _check(lib().enerf_psnr_accumulate(ptr(pred), ptr(gt), mask.data_ptr() if mask is not None else None,
                                       mask.element_size() if mask is not None else 0, n, acc.data_ptr(), stream()), "enerf_psnr_accumulate")


def pack_rgb8(rgb, H, W, out, flip_vertical=False):
    _check(lib().enerf_pack_rgb8(ptr(rgb), H, W, int(flip_vertical), out.data_ptr(), stream()), "enerf_pack_rgb8")
