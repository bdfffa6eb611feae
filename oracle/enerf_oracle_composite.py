"""TEST INFRASTRUCTURE -- CPU oracle for the layered ("composite") ENeRF forward.  NOT part of the product.

Restates lib/networks/enerf/network_composite.py (the `enerf_outdoor` configs: one or more
bbox-cropped foreground layers + one full-frame background layer, merged per pixel by a z-sort and
alpha compositing) on top of the building blocks of ``oracle/enerf_oracle.py``.  Same rules as that
module: torch fp32 primitives, every function cites the reference lines it follows (paths relative
to /root/reference), pinned against fixtures minted from the unmodified reference
(oracle/make_golden.py cases ``c5_*``), importable only by tests/, smoke() and bench.py's CPU arms.
"""
import torch
import torch.nn.functional as F

from . import enerf_oracle as O

BG_PLANES = (16, 4)   # network_composite.py:124: D=[16, 4][i] is hard-coded for the background volume


def nerf_novox(sd, p, f, viewdir_agg):
    """nerf_.NeRF.forward, nerf_.py:29-43: like nerf.NeRF but the voxel feature is NOT concatenated
    (nerf_.py:33-34), so lr0 is Linear(16,64) and color.0 is Linear(64+16+fc+4,64).
    f (B,N,S,fc+4) -> (B,N,4) = rgb(3), sigma(1)."""
    S = f.shape[2]
    vif = O.agg(sd, p + ".agg", f, viewdir_agg)
    x = F.relu(O._lin(sd, p + ".lr0.0", vif))
    sigma = F.softplus(O._lin(sd, p + ".sigma.0", x))
    xx = torch.cat([x, vif], dim=-1)[:, :, None].expand(-1, -1, S, -1)
    c = F.relu(O._lin(sd, p + ".color.2", F.relu(O._lin(sd, p + ".color.0", torch.cat([xx, f], dim=-1)))))
    w = F.softmax(c, dim=-2)
    rgb = (f[..., -7:-4] * w).sum(dim=-2)
    return torch.cat([rgb, sigma], dim=-1)


def _int_box(box, scale):
    """(bbox * scale).int() -> python ints (network_composite.py:88-89, utils.py:256-257,879-880)."""
    x, y, w, h = (box * scale).int()
    return x.item(), y.item(), w.item(), h.item()


def depth_values_composite(batch, cfg, level, D, inter, layer):
    """get_depth_values_composite, utils.py:153-214: get_depth_values with the layer's own near/far
    (batch['near_far'][:, layer], :170) at the first level and the layer's previous-level
    depth/std/near_far maps (:160-163) afterwards.  The arithmetic is get_depth_values' own."""
    b = dict(batch)
    b["near_far"] = batch["near_far"][:, layer]
    prev = (inter.get(f"depth_{level - 1}_{layer}"), inter.get(f"std_{level - 1}_{layer}"), inter.get(f"near_far_{level - 1}_{layer}"))
    return O.depth_values_for_level(b, cfg, level, D, *prev)


def homo_warp_crop(src_feat, proj, depth_values, xywh):
    """homo_warp_composite, utils.py:275-320: homo_warp restricted to the target window
    [y:y+h, x:x+w] of the (full-size) hypothesis maps -> (B,C,D,h,w)."""
    x, y, w, h = xywh
    B, D, Ht, Wt = depth_values.shape
    C, Hs, Ws = src_feat.shape[1:]
    dev = src_feat.device
    ys, xs = torch.meshgrid(torch.arange(Ht, dtype=torch.float32, device=dev), torch.arange(Wt, dtype=torch.float32, device=dev), indexing="ij")
    ys, xs = ys[y:y + h, x:x + w], xs[y:y + h, x:x + w]
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=dev)], 0)[None].expand(B, -1, -1).repeat(1, 1, D)
    q = proj[:, :, :3] @ pix + proj[:, :, 3:] / depth_values[:, :, y:y + h, x:x + w].reshape(B, 1, D * h * w)
    xy = q[:, :2] / torch.clamp_min(q[:, 2:], 1e-6)
    gx = xy[:, 0] / ((Ws - 1) / 2) - 1
    gy = xy[:, 1] / ((Hs - 1) / 2) - 1
    grid = torch.stack([gx, gy], -1).view(B, D, h * w, 2)
    out = F.grid_sample(src_feat, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.view(B, C, D, h, w)


def feature_volume_composite(feat, batch, cfg, level, D, inter, layer, xywh):
    """build_feature_volume_composite, utils.py:249-273 -> variance (B,C,D,hc,wc) over the window,
    depth_values (B,D,h,w) and near_far (B,2,h,w) over the full volume grid."""
    c = cfg.enerf.cas_config
    S = feat.shape[1]
    dv, nf = depth_values_composite(batch, cfg, level, D, inter, layer)
    pm = O.proj_mats(batch, c.im_feat_scale[level], c.volume_scale[level])
    s1, s2 = 0, 0
    for s in range(S):
        wv = homo_warp_crop(feat[:, s], pm[:, s], dv, xywh)
        s1 = s1 + wv
        s2 = s2 + wv ** 2
    return s2 / S - (s1 / S) ** 2, dv, nf      # utils.py:268


def build_rays_composite(depth, std, batch, cfg, near_far, level, box):
    """build_rays_composite, utils.py:216-247: build_rays on the full frame, then the rays inside the
    layer's window at render resolution, row-major -> (B, h*w, 12), [x,y,w,h]."""
    c = cfg.enerf.cas_config
    rays = O.build_rays(depth, std, batch, cfg, near_far, level)
    B = rays.shape[0]
    H, W = batch["src_inps"].shape[-2:]
    Hr, Wr = int(H * c.render_scale[level]), int(W * c.render_scale[level])
    x, y, w, h = _int_box(box, c.render_scale[level])
    rays = rays.reshape(B, Hr, Wr, rays.shape[-1])[:, y:y + h, x:x + w]
    return rays.reshape(B, -1, rays.shape[-1]), [x, y, w, h]


def render_rays_composite(sd, cfg, level, rays, batch, im_feat, src_inps, nerf_prefix):
    """Network.render_rays, network_composite.py:29-51 -> raw (B,N,Ns,4) and metric z (B,N,Ns).
    (get_vox_feat is evaluated by the reference, :45, but its result never reaches nerf_.NeRF.)"""
    c = cfg.enerf.cas_config
    ns = c.num_samples[level]
    xyz, _, z = O.sample_along_depth(rays, ns, c.depth_inv[level])
    B = xyz.shape[0]
    rgbs = O.unpreprocess(src_inps, c.render_scale[level])
    upf = c.render_scale[level] / c.im_ibr_scale[level]
    if upf != 1.0:
        b, s, ch, hh, ww = im_feat.shape
        im_feat = F.interpolate(im_feat.reshape(b * s, ch, hh, ww), None, scale_factor=upf, align_corners=True,
                                mode="bilinear").view(b, s, ch, int(hh * upf), int(ww * upf))
    ifd = O.img_feat(xyz, torch.cat([im_feat, rgbs], dim=2), batch, c.render_scale[level])
    raw = nerf_novox(sd, nerf_prefix, ifd, cfg.enerf.viewdir_agg).reshape(B, -1, ns, 4)
    return {"net_output": raw, "z_vals": 1.0 / z if c.depth_inv[level] else z}


def _batchify(sd, cfg, level, rays, batch, im_feat, src_inps, nerf_prefix):
    """batchify_rays, network_composite.py:53-63."""
    chunk = int(cfg.enerf.chunk_size)
    parts = [render_rays_composite(sd, cfg, level, rays[:, j:j + chunk], batch, im_feat, src_inps, nerf_prefix)
             for j in range(0, rays.shape[1], chunk)]
    return {k: torch.cat([p[k] for p in parts], dim=1) for k in parts[0]}


def parse_layer(layer, box, Hr, Wr, render_scale):
    """parse_layer, utils.py:875-887: the layer's samples placed into a zero full-frame canvas."""
    B, _, ns, _ = layer["net_output"].shape
    x, y, w, h = _int_box(box, render_scale)
    dev = layer["net_output"].device
    raw = torch.zeros(B, Hr, Wr, ns, 4, device=dev)
    z = torch.zeros(B, Hr, Wr, ns, device=dev)
    raw[:, y:y + h, x:x + w] = layer["net_output"].reshape(B, h, w, ns, 4)
    z[:, y:y + h, x:x + w] = layer["z_vals"].reshape(B, h, w, ns)
    return raw.reshape(B, -1, ns, 4), z.reshape(B, -1, ns)


def raw2outputs_composite(layers, batch, cfg, level, num_fg):
    """raw2outputs_composite, utils.py:889-942.  Foreground samples of all layers are concatenated and
    (only when num_fg > 1) sorted by z; the background samples are appended AFTER the sort, unsorted
    (:917-918); plain alpha compositing, depth = sum(w * z) (no softmax, unlike raw2outputs)."""
    c = cfg.enerf.cas_config
    H, W = batch["src_inps"].shape[-2:]
    Hr, Wr = int(H * c.render_scale[level]), int(W * c.render_scale[level])
    raw, z = parse_layer(layers[0], batch["bbox"][0][0], Hr, Wr, c.render_scale[level])
    for i in range(1, num_fg):
        r_, z_ = parse_layer(layers[i], batch["bbox"][0][i], Hr, Wr, c.render_scale[level])
        raw, z = torch.cat([raw, r_], dim=-2), torch.cat([z, z_], dim=-1)
    z_ori = z
    idx = None
    if num_fg > 1:
        z, idx = torch.sort(z, dim=-1)
        raw = raw.gather(dim=2, index=idx[..., None].repeat(1, 1, 1, 4))
    raw = torch.cat([raw, layers[-1]["net_output"]], dim=-2)
    z = torch.cat([z, layers[-1]["z_vals"]], dim=-1)
    alpha = 1.0 - torch.exp(-raw[..., 3])
    T = torch.cumprod(1.0 - alpha + 1e-10, dim=-1)[..., :-1]
    T = torch.cat([torch.ones_like(alpha[..., :1]), T], dim=-1)
    w = alpha * T
    rgb = torch.sum(w[..., None] * raw[..., :3], -2)
    depth = torch.sum(w * z, -1)
    # white_bkgd keeps its default False here: network_composite.py:142 does not pass it
    return {"rgb": rgb, "depth": depth, "weights": w, "net_output": raw, "idx": idx, "z_vals": z_ori}


def forward(sd, cfg, batch, intermediates=False):
    """Network.forward, network_composite.py:77-146."""
    c = cfg.enerf.cas_config
    L = int(cfg.num_fg_layers)
    B, S, _, H, W = batch["src_inps"].shape

    def feats_of(prefix):
        f2, f1, f0 = O.feature_net(sd, batch["src_inps"].reshape(B * S, 3, H, W), p=prefix)
        return {2: f0.reshape(B, S, -1, H, W), 1: f1.reshape(B, S, -1, H // 2, W // 2), 0: f2.reshape(B, S, -1, H // 4, W // 4)}

    feats, feats_bg = feats_of("feature_net"), feats_of("feature_net_bg")     # :78-79 (both on src_inps)
    ret, inter, mid = {}, {}, {}
    nf_all = batch["near_far"].clone()
    bg_batch = dict(batch)
    bg_batch["near_far"] = nf_all[:, -1]                                      # :118
    depth_ = std_ = nf_ = None
    for i in range(c.num):
        layers = []
        for l in range(L):
            xywh = _int_box(batch["bbox"][0][l], c.volume_scale[i])
            var, dv, nf = feature_volume_composite(feats[i], batch, cfg, i, c.volume_planes[i], inter, l, xywh)
            _, prob = O.cost_reg(sd, f"cost_reg_{i}_layer{l}", var, deep=False)    # MinCostRegNet at every level (:18)
            x, y, w, h = xywh
            Hv, Wv = dv.shape[-2:]
            prob = F.pad(prob, (x, Wv - x - w, y, Hv - y - h), "constant")         # :102
            depth, std = O.depth_regression(prob, dv, c.depth_inv[i])
            inter[f"depth_{i}_{l}"], inter[f"std_{i}_{l}"], inter[f"near_far_{i}_{l}"] = depth, std, nf
            mid.update({f"variance_{i}_{l}": var, f"depth_{i}_{l}": depth, f"std_{i}_{l}": std, f"near_far_{i}_{l}": nf})
            if c.render_if[i]:
                rays, _ = build_rays_composite(depth, std, batch, cfg, nf, i, batch["bbox"][0][l])
                layers.append(_batchify(sd, cfg, i, rays, batch, feats[c.render_im_feat_level[i]], batch["src_inps"],
                                        f"nerf_{i}_layer{l}"))
        var_, dv_, nf_ = O.feature_volume(feats_bg[i], bg_batch, cfg, i, depth_, std_, nf_, planes=BG_PLANES[i])
        _, prob_ = O.cost_reg(sd, f"cost_reg_{i}_bg", var_, deep=False)
        depth_, std_ = O.depth_regression(prob_, dv_, c.depth_inv[i])
        mid.update({f"depth_{i}_bg": depth_, f"std_{i}_bg": std_, f"near_far_{i}_bg": nf_})
        if c.render_if[i]:
            rays_ = O.build_rays(depth_, std_, bg_batch, cfg, nf_, i)
            layers.append(_batchify(sd, cfg, i, rays_, bg_batch, feats_bg[c.render_im_feat_level[i]], batch["bg_src_inps"],
                                    f"nerf_{i}_bg"))
            out = raw2outputs_composite(layers, batch, cfg, i, L)
            ret.update({f"{k}_level{i}": v for k, v in out.items()})
    return (ret, mid) if intermediates else ret
