"""Drop-in ``Network`` for the reference's plugin loader.

The reference builds its model with ``imp.load_source(cfg.network_module, cfg.network_path).Network()``
(/root/reference/lib/networks/make_network.py:5-9); pointing ``network_module`` at this file
(``network_module /root/repo/enerf_b200/network`` -- a dot-free value resolves to ``<value>.py``,
lib/config/config.py:166-168) swaps the render path for the B200 kernels and nothing else:

  * zero-argument constructor reading the global cfg                      (network.py:12-22)
  * identical parameter tree / state_dict keys / initialisers             (enerf_b200/params.py)
  * ``forward(batch) -> dict`` with the reference's output keys/shapes    (network.py:76-113)

``forward`` is inference-only (eval mode; the north-star path is render-time) and runs entirely on
the current CUDA stream with zero host synchronisation: ~30 launches of hand-written sm_100a
kernels through the C ABI in include/enerf_b200.h.  There is no CPU or PyTorch-op fallback.
"""
import os
import sys

import torch
import torch.nn as nn

if __package__ in (None, ""):  # loaded by file path through imp.load_source: make the package importable
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from enerf_b200 import capi, packing  # noqa: E402
from enerf_b200.config import get_cfg, snapshot  # noqa: E402
from enerf_b200.params import CostRegParams, FeatureParams, NerfParams  # noqa: E402


class Network(nn.Module):
    def __init__(self):
        super().__init__()
        cfg = get_cfg()
        cas = cfg.enerf.cas_config
        self.feature_net = FeatureParams()
        for i in range(cas.num):  # network.py:15-22
            setattr(self, f"cost_reg_{i}", CostRegParams(int(32 * (2 ** (-i))), deep=(i != 0)))
            setattr(self, f"nerf_{i}", NerfParams(cas.nerf_model_feat_ch[i] + 3, bool(cfg.enerf.viewdir_agg)))
        self._packed = None
        self._packed_key = None
        self._buffers_cache = {}
        # "tf32": MLP contractions on tcgen05 tensor cores (TF32 operands, fp32 accumulate) where the
        # tensor-core kernel is built for the configuration; "fp32": FP32-pipe kernels everywhere.
        self.precision = os.environ.get("ENERF_B200_PRECISION", "tf32")
        # two-stream schedule: the FeatureNet pyramid tail (laterals + smooth convs -> level_1/level_2
        # features, only needed from the level-1 cost volume on) runs on a side stream concurrently
        # with the level-0 chain (cost volume -> MinCostRegNet -> depth regression)
        self.overlap = os.environ.get("ENERF_B200_OVERLAP", "1") != "0"
        # network_human.py behaviour (render only rays inside batch['mask_at_box'] at the last level);
        # enerf_b200/network_human.py switches it on
        self.masked = False
        self._side = None
        # rows [r0, r1) of the render frame to generate rays for when the batch carries no rays_{i}
        # (None = the full frame); set by the ray-sharding renderer (enerf_b200/dist.py)
        self.ray_rows = None
        self.output_views = None   # {level: {"rgb","depth","weights"}} pre-allocated ray outputs (enerf_b200/dist.py)
        # Row-band layout of the LAST level (enerf_b200/dist.py, SURVEY.md section 8e option 1): with ray_rows set and
        # band_shard True, the cost volume / CostRegNet / regression of that level run only on the volume rows the
        # band's rays touch plus a halo that covers the regulariser's receptive field (+-30 rows for CostRegNet,
        # +-14 for MinCostRegNet, +1 for the bilinear / trilinear taps), so the band is bit-identical to the full frame.
        self.band_shard = False
        # masked path without the host read-back of the ray count (graph-capturable): outputs keep the reference's
        # keys, depth / weights are padded to the full ray count and `mask_count` (device int32) says how many are valid
        self.static_mask = False
        self.profile = False       # when True, CUDA events bracket every stage (see stage_times_ms)
        self._events = []

    # ---------------------------------------------------------------- packed weights (BN folded)
    def _fingerprint(self, want_feat):
        ver = 0
        first = None
        for t in self.state_dict(keep_vars=True).values():
            ver += t._version
            if first is None:
                first = (t.data_ptr(), t.device)
        return (ver, first, tuple(want_feat), self.precision)

    def packed_weights(self, levels):
        want_feat = [lv.render_if for lv in levels]
        key = self._fingerprint(want_feat)
        if self._packed is None or self._packed_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            dev = next(self.parameters()).device
            tcs = self.precision == "tf32"
            pk = {"feature": packing.pack_feature_net(sd, dev, tensor_cores=tcs)}
            for i, lv in enumerate(levels):
                pk[f"reg{i}"] = packing.pack_cost_reg(sd, f"cost_reg_{i}", int(32 * (2 ** (-i))), i != 0, dev, lv.render_if,
                                                      tensor_cores=tcs)
                vd = hasattr(getattr(self, f"nerf_{i}").agg, "view_fc")
                pk[f"nerf{i}"] = packing.pack_nerf(sd, f"nerf_{i}", lv.feat_ch + 3, vd, dev)
                if lv.feat_ch == 8:
                    pk[f"nerf_tc{i}"] = packing.pack_nerf_tc(sd, f"nerf_{i}", lv.feat_ch + 3, vd, dev)
            self._packed, self._packed_key = pk, key
        return self._packed

    def invalidate_packed(self):
        self._packed = None

    def train(self, mode=True):
        self._packed = None
        return super().train(mode)

    def _pack_img_feat(self, i, lv, feats, src, S, H, W, dev):
        """(S,Hr,Wr,feat_ch+4) records [features | un-preprocessed rgb | 0] the ray kernels gather from (csrc/feature_net.cu)."""
        Hr, Wr = int(H * lv.render_scale), int(W * lv.render_scale)
        imf = feats[lv.im_feat_level]
        if imf.shape[1] != Hr or imf.shape[2] != Wr or imf.shape[3] != lv.feat_ch:
            raise ValueError(f"level {i}: image features {tuple(imf.shape)} do not match render size {Hr}x{Wr}x{lv.feat_ch}")
        img = self._scratch(f"img{i}", S * Hr * Wr * (lv.feat_ch + 4), dev).view(S, Hr, Wr, lv.feat_ch + 4)
        capi.pack_img_feat(imf, src, img)
        return img

    # ---------------------------------------------------------------- scratch buffers (reused)
    def _scratch(self, name, numel, device, dtype=torch.float32):
        t = self._buffers_cache.get(name)
        if t is None or t.numel() < numel or t.device != device or t.dtype != dtype:
            t = torch.empty(numel, device=device, dtype=dtype)
            self._buffers_cache[name] = t
        return t[:numel]

    # ---------------------------------------------------------------- row-band layout (enerf_b200/dist.py)
    HALO = {True: 32, False: 16}     # CostRegNet: receptive field +-30 rows; MinCostRegNet: +-14

    def _band_rows(self, lv, H, h, deep):
        """Volume rows [y0, y1) a rank must regularise so that its ray rows ``self.ray_rows`` come out exactly."""
        Hr = int(H * lv.render_scale)
        r0, r1 = self.ray_rows
        div = 8 if deep else 4
        if (r0 * h) % Hr or (r1 * h) % Hr or ((r0 * h) // Hr) % div or ((r1 * h) // Hr) % div:
            raise ValueError(f"band rows {r0}:{r1} of {Hr} do not map onto volume rows aligned to the regulariser's stride {div}")
        v0, v1 = r0 * h // Hr, r1 * h // Hr
        halo = self.HALO[deep]
        return max(0, v0 - halo), min(h, v1 + halo)

    # ---------------------------------------------------------------- per-stage CUDA-event timing
    def _mark(self, name):
        if self.profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._events.append((name, ev))

    def stage_times_ms(self):
        """{stage: ms} of the last profiled forward (call after torch.cuda.synchronize())."""
        out = {}
        for (n0, e0), (n1, e1) in zip(self._events[:-1], self._events[1:]):
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out

    # ---------------------------------------------------------------- forward
    def forward(self, batch):
        if self.training:
            raise NotImplementedError("enerf_b200.Network implements the render-time (eval) path only; call .eval()")
        src = batch["src_inps"]
        if not src.is_cuda:
            raise ValueError("enerf_b200.Network runs on CUDA tensors only (no CPU fallback); move the batch to the GPU")
        B = src.shape[0]
        if B == 1:
            return self._forward_one(batch, 0)
        outs = [self._forward_one(batch, b) for b in range(B)]
        return {k: torch.cat([o[k] for o in outs], dim=0) for k in outs[0]}

    def _forward_one(self, batch, b):
        cfg = get_cfg()
        levels = snapshot(cfg)
        src = batch["src_inps"][b].float().contiguous()            # (S,3,H,W)
        S, _, H, W = src.shape
        dev = src.device
        if S < 2 or S > capi.MAX_VIEWS:
            raise ValueError(f"need 2..{capi.MAX_VIEWS} source views, got {S}")
        if len(levels) > capi.MAX_LEVELS:
            raise NotImplementedError(f"at most {capi.MAX_LEVELS} cascade levels")
        pk = self.packed_weights(levels)
        f32 = dict(device=dev, dtype=torch.float32)
        self._events = []
        self._mark("start")

        cam = self._scratch("cam", capi.CAM_FLOATS, dev)
        scales = [(lv.im_feat_scale, lv.volume_scale, lv.render_scale) for lv in levels]
        capi.camera_setup(batch["src_exts"][b].float().contiguous(), batch["src_ixts"][b].float().contiguous(),
                          batch["tar_ext"][b].float().contiguous(), batch["tar_ixt"][b].float().contiguous(),
                          batch["near_far"][b].float().contiguous(), scales, cam)

        if H % 4 or W % 4:
            raise ValueError(f"H={H}, W={W} must be multiples of 4 (FeatureNet strides)")
        feats = {0: torch.empty((S, H // 4, W // 4, 32), **f32), 1: torch.empty((S, H // 2, W // 2, 16), **f32),
                 2: torch.empty((S, H, W, 8), **f32)}
        ws = self._scratch("feat_ws", capi.feature_net_workspace_bytes(S, H, W) // 4, dev)
        self._mark("camera_setup")
        tcs = self.precision == "tf32"
        fork = self.overlap and not self.profile and len(levels) > 1
        tail_done = None
        prepacked = {}
        # a render level that gathers from the full-resolution level-2 features: FeatureNet writes its (feature | rgb) records
        # itself (the fused lat0 + smooth0 launch's epilogue on the tensor-core path), no separate pack pass
        img2, img2_level = None, None
        for i, lv in enumerate(levels):
            if (lv.render_if and lv.im_feat_level == 2 and lv.feat_ch == 8 and lv.render_scale == lv.im_ibr_scale
                    and int(H * lv.render_scale) == H and int(W * lv.render_scale) == W):
                img2 = self._scratch(f"img{i}", S * H * W * 12, dev).view(S, H, W, 12)
                img2_level = i
                break
        if fork:
            main = torch.cuda.current_stream()
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            capi.feature_net(pk["feature"], src, feats[0], feats[1], feats[2], ws, tensor_cores=tcs, part=1)
            trunk_done = torch.cuda.Event()
            trunk_done.record(main)
            with torch.cuda.stream(self._side):
                self._side.wait_event(trunk_done)
                capi.feature_net(pk["feature"], src, feats[0], feats[1], feats[2], ws, tensor_cores=tcs, part=2, img_feat_rgb=img2)
                if img2 is not None:
                    prepacked[img2_level] = img2
                # the (feature | rgb) records of the other render levels depend on the pyramid and the source images only: pack them
                # here, next to the level-0 chain, instead of serially in front of the ray launch
                for i, lv in enumerate(levels):
                    if lv.render_if and lv.render_scale == lv.im_ibr_scale and i not in prepacked:
                        prepacked[i] = self._pack_img_feat(i, lv, feats, src, S, H, W, dev)
                tail_done = torch.cuda.Event()
                tail_done.record(self._side)
            for tns in (feats[1], feats[2], ws, src):
                tns.record_stream(self._side)
        else:
            capi.feature_net(pk["feature"], src, feats[0], feats[1], feats[2], ws, tensor_cores=tcs, img_feat_rgb=img2)
            if img2 is not None:
                prepacked[img2_level] = img2
        self._mark("feature_net")

        ret = {}
        depth = std = nf = None
        for i, lv in enumerate(levels):
            h, w, D = int(H * lv.volume_scale), int(W * lv.volume_scale), lv.planes
            deep = i != 0
            div = 8 if deep else 4
            if D % div or h % div or w % div:
                raise ValueError(f"level {i}: volume {D}x{h}x{w} must be divisible by {div} (U-Net skip connections)")
            if i > 0 and not lv.prev_depth_inv:
                raise NotImplementedError("cascade needs depth_inv on the previous level (reference: utils.py:129-130)")
            if i > 0 and tail_done is not None:     # join: level >= 1 reads the pyramid features
                torch.cuda.current_stream().wait_event(tail_done)
                tail_done = None
            feat = feats[i]
            ends = torch.empty((2, h, w), **f32)
            nf_new = torch.empty((2, h, w), **f32)
            capi.depth_hypotheses(cam, depth, std, nf, h, w, D, lv.depth_inv, ends, nf_new)
            nf = nf_new
            # rows [y0, y1) of the volume this call regularises: everything, or the rank's band + halo
            y0, y1 = 0, h
            if self.band_shard and self.ray_rows is not None and i == len(levels) - 1 and lv.render_if:
                y0, y1 = self._band_rows(lv, H, h, deep)
            hc = y1 - y0
            var = self._scratch(f"var{i}", D * hc * w * feat.shape[-1], dev).view(D, hc, w, feat.shape[-1])
            self._mark(f"depth_hypotheses_{i}")
            if hc == h:
                capi.cost_volume(cam, i, feat, ends, D, h, w, lv.depth_inv, var)
            else:
                capi.cost_volume_window(cam, i, feat, ends, D, h, w, [0, y0, w, hc], lv.depth_inv, var)
            self._mark(f"cost_volume_{i}")
            vol = torch.empty((D, hc, w, 8), **f32) if lv.render_if else None
            prob = self._scratch(f"prob{i}", D * hc * w, dev).view(D, hc, w)
            rws = self._scratch(f"reg_ws{i}", capi.cost_reg_workspace_bytes(deep, D, hc, w) // 4, dev)
            capi.cost_reg(pk[f"reg{i}"], deep, var, vol, prob, rws, tensor_cores=tcs)
            self._mark(f"cost_reg_{i}")
            lvl_views = (self.output_views or {}).get(i) or {}
            depth = torch.empty((h, w), **f32)
            std = lvl_views["std"] if "std" in lvl_views else torch.empty((h, w), **f32)
            mvs = lvl_views["depth_mvs"] if "depth_mvs" in lvl_views else torch.empty((h, w), **f32)
            if std.shape != (h, w) or mvs.shape != (h, w):
                raise ValueError("output_views std / depth_mvs shapes do not match the level's volume resolution")
            if hc == h:
                capi.depth_regress(prob, ends, lv.depth_inv, depth, std, mvs)
            else:   # rows outside [y0, y1) are NOT this rank's (they come out as the uniform-probability value)
                capi.depth_regress_window(prob, [0, y0, w, hc], ends, lv.depth_inv, depth, std, mvs)
            self._mark(f"depth_regress_{i}")
            if not lv.render_if:
                continue
            if lv.render_scale != lv.im_ibr_scale:
                raise NotImplementedError("render_scale != im_ibr_scale (feature up-sampling in render_rays, network.py:30-32) "
                                          "is not used by any shipped config and is not implemented")
            Hr, Wr = int(H * lv.render_scale), int(W * lv.render_scale)
            if tail_done is not None and (lv.im_feat_level > 0 or i in prepacked):
                torch.cuda.current_stream().wait_event(tail_done)
                tail_done = None
            img = prepacked[i] if i in prepacked else self._pack_img_feat(i, lv, feats, src, S, H, W, dev)
            self._mark(f"pack_img_feat_{i}")
            if f"rays_{i}" in batch:
                rays = batch[f"rays_{i}"][b].float().contiguous()
            else:   # no rays from the data layer: full-frame rays generated on device
                r0, r1 = self.ray_rows if self.ray_rows is not None else (0, Hr)
                rays = torch.empty(((r1 - r0) * Wr, 8), **f32)
                capi.generate_rays(batch["tar_ext"][b].float().contiguous(), batch["tar_ixt"][b].float().contiguous(),
                                   lv.render_scale, Wr, r0, r1 - r0, rays)
            n_full = rays.shape[0]
            use_mask = self.masked and "mask_at_box" in batch and i == len(levels) - 1
            if use_mask:   # rays = rays[mask_at_box]  (network_human.py:90-92), order-preserving, on device
                mask = batch["mask_at_box"][b].reshape(-1).contiguous()
                if mask.numel() != n_full:
                    raise ValueError(f"mask_at_box has {mask.numel()} elements, the level renders {n_full} rays")
                midx = torch.empty(n_full, device=dev, dtype=torch.int32)
                rays_c = torch.empty((n_full, 8), **f32)
                cnt = torch.empty(1, device=dev, dtype=torch.int32)
                mws = self._scratch("mask_ws", capi.mask_compact_workspace_bytes(n_full) // 4 + 1, dev)
                capi.mask_compact(mask, rays, midx, rays_c, cnt, mws)
                if self.static_mask:   # no read-back: launches sized for n_full, the kernels stop at the device-side count
                    rays = rays_c
                else:                  # the output shapes depend on the count: one D2H read, as the reference's boolean indexing implies
                    n_sel = int(cnt.item())
                    rays = rays_c[:n_sel]
            n_dev = cnt if (use_mask and self.static_mask) else None
            N = rays.shape[0]
            views = (self.output_views or {}).get(i)
            if views is not None and "rgb" in views:   # write straight into the caller's (gather) buffer
                rgb, dmap, wts = views["rgb"], views["depth"], views["weights"]
                if rgb.shape != (N, 3) or dmap.shape != (N,) or wts.shape != (N, lv.num_samples):
                    raise ValueError("output_views shapes do not match the ray batch")
            else:
                alloc = torch.zeros if n_dev is not None else torch.empty     # static mask: rows past the count stay zero
                rgb = alloc((N, 3), **f32)
                dmap = alloc((N,), **f32)
                wts = alloc((N, lv.num_samples), **f32)
            if N == 0:
                pass
            elif self.precision == "tf32" and capi.tc_ray_kernel_supports(lv.feat_ch, S, lv.num_samples):
                capi.render_rays_tc(cam, i, pk[f"nerf_tc{i}"], rays, depth, std, nf, vol, img, lv.feat_ch, lv.num_samples,
                                    lv.depth_inv, bool(cfg.enerf.white_bkgd), bool(cfg.enerf.viewdir_agg), rgb, dmap, wts,
                                    vol_row0=y0, n_rays_dev=n_dev)
            else:
                capi.render_rays(cam, i, pk[f"nerf{i}"], rays, depth, std, nf, vol, img, lv.feat_ch, lv.num_samples, lv.depth_inv,
                                 bool(cfg.enerf.white_bkgd), bool(cfg.enerf.viewdir_agg), rgb, dmap, wts, vol_row0=y0, n_rays_dev=n_dev)
            self._mark(f"render_rays_{i}")
            if use_mask:   # rgb scattered into a zero image, depth / weights stay compact (network_human.py:102-106)
                rgb_full = torch.zeros((n_full, 3), **f32)
                if n_dev is not None:
                    capi.scatter_rows(rgb, midx, N, rgb_full, m_dev=n_dev)
                    ret["mask_count"] = cnt
                elif N > 1:
                    capi.scatter_rows(rgb, midx, N, rgb_full)
                rgb = rgb_full
            ret.update({f"rgb_level{i}": rgb[None], f"depth_level{i}": dmap[None], f"weights_level{i}": wts[None],
                        f"depth_mvs_level{i}": mvs[None], f"std_level{i}": std[None]})
        if tail_done is not None:                   # always join the side stream before returning
            torch.cuda.current_stream().wait_event(tail_done)
        return ret
