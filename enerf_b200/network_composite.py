"""Drop-in ``Network`` for the reference's layered scenes (``network_module`` of
configs/enerf/enerf_outdoor/*.yaml -> /root/reference/lib/networks/enerf/network_composite.py).

``cfg.num_fg_layers`` foreground layers, each with its own bbox window (``batch['bbox']``), near/far
(``batch['near_far'][:, l]``), MinCostRegNet and nerf_.NeRF, plus a full-frame background layer that
sees the background plates (``batch['bg_src_inps']``) and has its own FeatureNet; the layers' samples
are merged per pixel (z-sort across foreground layers, background behind) and alpha-composited.

Same contract as ``enerf_b200.network.Network``: zero-argument constructor reading the global cfg,
the reference's parameter tree / state_dict keys (network_composite.py:12-27), ``forward(batch)``
returning the reference's keys and shapes (:77-146), inference only, CUDA only, no fallback.

What runs per level (all through the C ABI, include/enerf_b200.h "Layered rendering"):
  per foreground layer: depth_hypotheses_layer -> cost_volume_window (bbox at volume resolution) ->
  cost_reg (depth head only: nerf_.NeRF never reads the feature volume) -> depth_regress_window ->
  render_rays_raw over the bbox at render resolution;
  background: depth_hypotheses -> cost_volume -> cost_reg -> depth_regress -> render_rays_raw (full frame);
  composite_layers.
The reference's zero-padded canvases (F.pad of the probability volume, parse_layer's (H,W,Ns,4)
buffers), the sort/gather passes and get_vox_feat (computed, then discarded by nerf_) do not exist here.
"""
import os
import sys

import torch
import torch.nn as nn

if __package__ in (None, ""):  # loaded by file path through imp.load_source
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from enerf_b200 import capi, packing  # noqa: E402
from enerf_b200.config import get_cfg, snapshot  # noqa: E402
from enerf_b200.params import CostRegParams, FeatureParams, NerfParams  # noqa: E402

BG_PLANES = (16, 4)   # network_composite.py:124 hard-codes the background volume depths


def _int_box(box, scale):
    """(bbox * scale).int() with the reference's float32 arithmetic (network_composite.py:88, utils.py:256,879)."""
    return [int(v) for v in (box * scale).int().tolist()]


class Network(nn.Module):
    def __init__(self):
        super().__init__()
        cfg = get_cfg()
        cas = cfg.enerf.cas_config
        self.num_fg_layers = int(cfg.num_fg_layers)
        if self.num_fg_layers < 1:
            raise ValueError("network_composite needs cfg.num_fg_layers >= 1")
        vd = bool(cfg.enerf.viewdir_agg)
        self.feature_net = FeatureParams()
        self.feature_net_bg = FeatureParams()
        for i in range(cas.num):   # network_composite.py:16-25 (MinCostRegNet and nerf_.NeRF at every level)
            for l in range(self.num_fg_layers):
                setattr(self, f"cost_reg_{i}_layer{l}", CostRegParams(int(32 * (2 ** (-i))), deep=False))
                setattr(self, f"nerf_{i}_layer{l}", NerfParams(cas.nerf_model_feat_ch[i] + 3, vd, vox_ch=0))
            setattr(self, f"cost_reg_{i}_bg", CostRegParams(int(32 * (2 ** (-i))), deep=False))
            setattr(self, f"nerf_{i}_bg", NerfParams(cas.nerf_model_feat_ch[i] + 3, vd, vox_ch=0))
        self.precision = os.environ.get("ENERF_B200_PRECISION", "tf32")   # conv stacks on tcgen05 ("tf32") or FP32 pipes
        # the background chain (its FeatureNet included) runs on a side stream next to the foreground layers
        self.overlap = os.environ.get("ENERF_B200_OVERLAP", "1") != "0"
        self._packed = None
        self._packed_key = None
        self._buffers_cache = {}
        self._side = None

    # ---------------------------------------------------------------- packed weights (BN folded)
    def _fingerprint(self):
        ver, first = 0, None
        for t in self.state_dict(keep_vars=True).values():
            ver += t._version
            if first is None:
                first = (t.data_ptr(), t.device)
        return (ver, first, self.precision)

    def packed_weights(self, levels):
        key = self._fingerprint()
        if self._packed is None or self._packed_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            dev = next(self.parameters()).device
            tcs = self.precision == "tf32"
            vd = hasattr(getattr(self, "nerf_0_bg").agg, "view_fc")
            pk = {"feature": packing.pack_feature_net(sd, dev, p="feature_net", tensor_cores=tcs),
                  "feature_bg": packing.pack_feature_net(sd, dev, p="feature_net_bg", tensor_cores=tcs)}
            for i, lv in enumerate(levels):
                for tag in [f"layer{l}" for l in range(self.num_fg_layers)] + ["bg"]:
                    pk[f"reg{i}_{tag}"] = packing.pack_cost_reg(sd, f"cost_reg_{i}_{tag}", int(32 * (2 ** (-i))), False, dev, False,
                                                                tensor_cores=tcs)
                    pk[f"nerf{i}_{tag}"] = packing.pack_nerf_novox(sd, f"nerf_{i}_{tag}", lv.feat_ch + 3, vd, dev)
                    if lv.feat_ch == 8:
                        pk[f"nerf_tc{i}_{tag}"] = packing.pack_nerf_tc_novox(sd, f"nerf_{i}_{tag}", lv.feat_ch + 3, vd, dev)
            self._packed, self._packed_key = pk, key
        return self._packed

    def train(self, mode=True):
        self._packed = None
        return super().train(mode)

    def _scratch(self, name, numel, device, dtype=torch.float32):
        t = self._buffers_cache.get(name)
        if t is None or t.numel() < numel or t.device != device or t.dtype != dtype:
            t = torch.empty(numel, device=device, dtype=dtype)
            self._buffers_cache[name] = t
        return t[:numel]

    # ---------------------------------------------------------------- forward
    def forward(self, batch):
        if self.training:
            raise NotImplementedError("enerf_b200 network_composite implements the render-time (eval) path only; call .eval()")
        src = batch["src_inps"]
        if not src.is_cuda:
            raise ValueError("enerf_b200 runs on CUDA tensors only (no CPU fallback); move the batch to the GPU")
        outs = [self._forward_one(batch, b) for b in range(src.shape[0])]
        if len(outs) == 1:
            return outs[0]
        return {k: (None if outs[0][k] is None else torch.cat([o[k] for o in outs], dim=0)) for k in outs[0]}

    def _feature_net(self, pk, src, tag, tcs):
        S, _, H, W = src.shape
        f32 = dict(device=src.device, dtype=torch.float32)
        feats = {0: torch.empty((S, H // 4, W // 4, 32), **f32), 1: torch.empty((S, H // 2, W // 2, 16), **f32),
                 2: torch.empty((S, H, W, 8), **f32)}
        ws = self._scratch(f"feat_ws_{tag}", capi.feature_net_workspace_bytes(S, H, W) // 4, src.device)
        capi.feature_net(pk, src, feats[0], feats[1], feats[2], ws, tensor_cores=tcs)
        return feats

    def _volume_stage(self, pk_reg, tag, i, lv, D, hw, cam, feat, first_nf, prev, window, tcs):
        """hypotheses -> (windowed) cost volume -> MinCostRegNet depth head -> depth regression.
        Returns (depth, std, near_far) over the full volume grid."""
        dev = feat.device
        f32 = dict(device=dev, dtype=torch.float32)
        h, w = hw
        ends, nf = torch.empty((2, h, w), **f32), torch.empty((2, h, w), **f32)
        if first_nf is None:
            capi.depth_hypotheses(cam, *prev, h, w, D, lv.depth_inv, ends, nf)
        else:
            capi.depth_hypotheses_layer(first_nf, *prev, h, w, D, lv.depth_inv, ends, nf)
        x, y, wc, hc = window if window is not None else (0, 0, w, h)
        if D % 4 or hc % 4 or wc % 4:
            raise ValueError(f"level {i} {tag}: volume {D}x{hc}x{wc} must be divisible by 4 (MinCostRegNet skip connections)")
        C = feat.shape[-1]
        var = self._scratch(f"var_{tag}_{i}", D * hc * wc * C, dev).view(D, hc, wc, C)
        if window is None:
            capi.cost_volume(cam, i, feat, ends, D, h, w, lv.depth_inv, var)
        else:
            capi.cost_volume_window(cam, i, feat, ends, D, h, w, window, lv.depth_inv, var)
        prob = self._scratch(f"prob_{tag}_{i}", D * hc * wc, dev).view(D, hc, wc)
        rws = self._scratch(f"reg_ws_{tag}_{i}", capi.cost_reg_workspace_bytes(False, D, hc, wc) // 4, dev)
        capi.cost_reg(pk_reg, False, var, None, prob, rws, tensor_cores=tcs)
        depth, std = torch.empty((h, w), **f32), torch.empty((h, w), **f32)
        if window is None:
            capi.depth_regress(prob, ends, lv.depth_inv, depth, std, None)
        else:
            capi.depth_regress_window(prob, window, ends, lv.depth_inv, depth, std)
        return depth, std, nf

    def _raw_rays(self, pk, tag, S, cam, i, rays, window, depth, std, nf, img, lv, vd, raw, zbuf, off):
        """Per-sample (rgb, sigma, z) of the rays in `window`: tcgen05 MLP where the tensor-core kernel is
        built for the configuration and precision allows, else the FP32-pipe kernel."""
        if self.precision == "tf32" and capi.tc_ray_kernel_supports(lv.feat_ch, S, lv.num_samples):
            capi.render_rays_raw_tc(cam, i, pk[f"nerf_tc{tag}"], rays, window, depth, std, nf, img, lv.feat_ch, lv.num_samples,
                                    lv.depth_inv, vd, raw, zbuf, off)
        else:
            capi.render_rays_raw(cam, i, pk[f"nerf{tag}"], rays, window, depth, std, nf, img, lv.feat_ch, lv.num_samples, lv.depth_inv,
                                 vd, raw, zbuf, off)

    def _forward_one(self, batch, b):
        cfg = get_cfg()
        levels = snapshot(cfg)
        L = self.num_fg_layers
        src = batch["src_inps"][b].float().contiguous()            # (S,3,H,W)
        bg_src = batch["bg_src_inps"][b].float().contiguous()
        S, _, H, W = src.shape
        dev = src.device
        if S < 2 or S > capi.MAX_VIEWS:
            raise ValueError(f"need 2..{capi.MAX_VIEWS} source views, got {S}")
        if H % 4 or W % 4:
            raise ValueError(f"H={H}, W={W} must be multiples of 4 (FeatureNet strides)")
        near_far = batch["near_far"][b].float().contiguous()       # (L+1, 2): foreground layers, background last
        boxes = batch["bbox"][b].detach().float().cpu()            # launch geometry: read once (the reference .item()s 4 values per layer and level)
        if near_far.shape[0] != L + 1 or boxes.shape[0] < L:
            raise ValueError(f"need {L} bboxes and {L + 1} near_far rows, got {tuple(boxes.shape)} / {tuple(near_far.shape)}")
        pk = self.packed_weights(levels)
        tcs = self.precision == "tf32"
        vd = bool(cfg.enerf.viewdir_agg)
        f32 = dict(device=dev, dtype=torch.float32)

        cam = self._scratch("cam", capi.CAM_FLOATS, dev)
        scales = [(lv.im_feat_scale, lv.volume_scale, lv.render_scale) for lv in levels]
        capi.camera_setup(batch["src_exts"][b].float().contiguous(), batch["src_ixts"][b].float().contiguous(),
                          batch["tar_ext"][b].float().contiguous(), batch["tar_ixt"][b].float().contiguous(),
                          near_far[L].contiguous(), scales, cam)       # cam.near_far = the background's (:118)

        main = torch.cuda.current_stream()
        side = main
        if self.overlap:
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            side = self._side
            side.wait_stream(main)
        feats = self._feature_net(pk["feature"], src, "fg", tcs)
        with torch.cuda.stream(side):
            feats_bg = self._feature_net(pk["feature_bg"], src, "bg", tcs)   # :79: the background net also sees src_inps

        ret = {}
        prev_fg = [(None, None, None)] * L
        prev_bg = (None, None, None)
        for i, lv in enumerate(levels):
            if i > 0 and not lv.prev_depth_inv:
                raise NotImplementedError("cascade needs depth_inv on the previous level (reference: utils.py:197-198)")
            if lv.render_if and lv.render_scale != lv.im_ibr_scale:
                raise NotImplementedError("render_scale != im_ibr_scale is not used by any shipped config and is not implemented")
            Hr, Wr = int(H * lv.render_scale), int(W * lv.render_scale)
            hw = (int(H * lv.volume_scale), int(W * lv.volume_scale))
            ns = lv.num_samples
            n_fg, n_tot = L * ns, L * ns + ns
            if lv.render_if:
                raw = torch.empty((Hr * Wr, n_tot, 4), **f32)
                zbuf = torch.empty((Hr * Wr, n_tot), **f32)
                if f"rays_{i}" in batch:
                    rays = batch[f"rays_{i}"][b].float().contiguous()
                else:
                    rays = torch.empty((Hr * Wr, 8), **f32)
                    capi.generate_rays(batch["tar_ext"][b].float().contiguous(), batch["tar_ixt"][b].float().contiguous(),
                                       lv.render_scale, Wr, 0, Hr, rays)
                if rays.shape[0] != Hr * Wr:
                    raise ValueError(f"level {i}: rays_{i} must hold the full {Hr}x{Wr} frame")
                imf = feats[lv.im_feat_level]
                img = self._scratch(f"img{i}", S * Hr * Wr * (lv.feat_ch + 4), dev).view(S, Hr, Wr, lv.feat_ch + 4)
                capi.pack_img_feat(imf, src, img)
            r_boxes = []
            for l in range(L):
                win_v = _int_box(boxes[l], lv.volume_scale)
                depth, std, nf = self._volume_stage(pk[f"reg{i}_layer{l}"], f"l{l}", i, lv, lv.planes, hw, cam, feats[i], near_far[l],
                                                    prev_fg[l], win_v, tcs)
                prev_fg[l] = (depth, std, nf)
                if lv.render_if:
                    win_r = _int_box(boxes[l], lv.render_scale)
                    r_boxes.append(win_r)
                    self._raw_rays(pk, f"{i}_layer{l}", S, cam, i, rays, win_r, depth, std, nf, img, lv, vd, raw, zbuf, l * ns)
            with torch.cuda.stream(side):
                depth_, std_, nf_ = self._volume_stage(pk[f"reg{i}_bg"], "bg", i, lv, BG_PLANES[i], hw, cam, feats_bg[i], None, prev_bg,
                                                       None, tcs)
                prev_bg = (depth_, std_, nf_)
                if lv.render_if:
                    if side is not main:
                        side.wait_stream(main)      # rays / raw / zbuf were produced on the main stream
                    img_bg = self._scratch(f"img_bg{i}", S * Hr * Wr * (lv.feat_ch + 4), dev).view(S, Hr, Wr, lv.feat_ch + 4)
                    capi.pack_img_feat(feats_bg[lv.im_feat_level], bg_src, img_bg)
                    self._raw_rays(pk, f"{i}_bg", S, cam, i, rays, [0, 0, Wr, Hr], depth_, std_, nf_, img_bg, lv, vd, raw, zbuf, n_fg)
            if not lv.render_if:
                continue
            if side is not main:
                main.wait_stream(side)
            rgb = torch.empty((Hr * Wr, 3), **f32)
            dmap = torch.empty((Hr * Wr,), **f32)
            wts = torch.empty((Hr * Wr, n_tot), **f32)
            net_out = torch.empty((Hr * Wr, n_tot, 4), **f32)
            z_vals = torch.empty((Hr * Wr, n_fg), **f32)
            idx = torch.empty((Hr * Wr, n_fg), device=dev, dtype=torch.int64) if L > 1 else None
            capi.composite_layers(raw, zbuf, Hr, Wr, L, ns, ns, r_boxes, rgb, dmap, wts, net_out, idx, z_vals)
            ret.update({f"rgb_level{i}": rgb[None], f"depth_level{i}": dmap[None], f"weights_level{i}": wts[None],
                        f"net_output_level{i}": net_out[None], f"idx_level{i}": None if idx is None else idx[None],
                        f"z_vals_level{i}": z_vals[None]})
        if side is not main:
            main.wait_stream(side)     # every side-stream access is ordered before whatever the caller enqueues next
        return ret
