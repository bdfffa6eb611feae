"""TEST INFRASTRUCTURE -- CPU oracle for the ENeRF render-time hot path.  NOT part of the product.

A functional torch-fp32 restatement of the reference algorithm (zju3dv/ENeRF,
lib/networks/enerf/{network,feature_net,cost_reg_net,nerf,utils}.py), written against a plain
``state_dict`` + cfg instead of nn.Modules.  The arithmetic primitives (conv, grid_sample,
interpolate, softmax, ...) are torch's -- the same third-party library the reference calls
(all of the reference's arithmetic lives in torch; README.md:20 pins 1.9.0,
this image has 2.11 and the op semantics relied on are unchanged).

Pinning: the reference's own tests hold NO golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, run in the authoring container by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (see tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product (enerf_b200/) never does.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm default, used by ConvBnReLU{,3D} (lib/networks/enerf/utils.py:10-33)


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def _bn(sd, p, x):
    """eval-mode batch norm with running statistics (utils.py:17,30: norm_act(out_channels))."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=BN_EPS)


def _cbr2d(sd, p, x, stride, pad):
    """ConvBnReLU.forward, utils.py:19-20 (conv has no bias)."""
    return F.relu(_bn(sd, p + ".bn", F.conv2d(x, sd[p + ".conv.weight"], None, stride, pad)))


def _cbr3d(sd, p, x, stride=1):
    """ConvBnReLU3D.forward, utils.py:32-33 (k=3, pad=1, no bias)."""
    return F.relu(_bn(sd, p + ".bn", F.conv3d(x, sd[p + ".conv.weight"], None, stride, 1)))


def _deconv_bn(sd, p, x):
    """ConvTranspose3d(k3,s2,p1,output_padding 1,bias False)+BN, no ReLU (cost_reg_net.py:19-33)."""
    y = F.conv_transpose3d(x, sd[p + ".0.weight"], None, stride=2, padding=1, output_padding=1)
    return _bn(sd, p + ".1", y)


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def feature_net(sd, x, p="feature_net"):
    """FeatureNet.forward, feature_net.py:27-36.  x (N,3,H,W) -> feat2 (N,32,H/4,W/4),
    feat1 (N,16,H/2,W/2), feat0 (N,8,H,W)."""
    c0 = _cbr2d(sd, p + ".conv0.1", _cbr2d(sd, p + ".conv0.0", x, 1, 1), 1, 1)
    c1 = _cbr2d(sd, p + ".conv1.1", _cbr2d(sd, p + ".conv1.0", c0, 2, 2), 1, 1)
    c2 = _cbr2d(sd, p + ".conv2.1", _cbr2d(sd, p + ".conv2.0", c1, 2, 2), 1, 1)
    feat2 = F.conv2d(c2, sd[p + ".toplayer.weight"], sd[p + ".toplayer.bias"])
    f1 = _up2(feat2) + F.conv2d(c1, sd[p + ".lat1.weight"], sd[p + ".lat1.bias"])
    f0 = _up2(f1) + F.conv2d(c0, sd[p + ".lat0.weight"], sd[p + ".lat0.bias"])
    feat1 = F.conv2d(f1, sd[p + ".smooth1.weight"], sd[p + ".smooth1.bias"], padding=1)
    feat0 = F.conv2d(f0, sd[p + ".smooth0.weight"], sd[p + ".smooth0.bias"], padding=1)
    return feat2, feat1, feat0


def cost_reg(sd, p, x, deep):
    """CostRegNet.forward (deep=True, cost_reg_net.py:35-48) / MinCostRegNet.forward (:75-86).
    x (B,C,D,h,w) -> (feat (B,8,D,h,w), depth_prob (B,D,h,w))."""
    c0 = _cbr3d(sd, p + ".conv0", x)
    c2 = _cbr3d(sd, p + ".conv2", _cbr3d(sd, p + ".conv1", c0, 2))
    c4 = _cbr3d(sd, p + ".conv4", _cbr3d(sd, p + ".conv3", c2, 2))
    y = c4
    if deep:
        y = _cbr3d(sd, p + ".conv6", _cbr3d(sd, p + ".conv5", c4, 2))
        y = c4 + _deconv_bn(sd, p + ".conv7", y)
    y = c2 + _deconv_bn(sd, p + ".conv9", y)
    y = c0 + _deconv_bn(sd, p + ".conv11", y)
    feat = F.conv3d(y, sd[p + ".feat_conv.0.weight"], None, 1, 1)
    prob = F.conv3d(y, sd[p + ".depth_conv.0.weight"], None, 1, 1)
    return feat, prob[:, 0]


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def agg(sd, p, f, viewdir_agg=True):
    """Agg.forward, nerf.py:74-89.  f (B,N,S,fc+4) -> (B,N,16)."""
    if viewdir_agg:
        g = f[..., :-4] + F.relu(_lin(sd, p + ".view_fc.0", f[..., -4:]))
    else:
        g = f[..., :-4]
    S = g.shape[-2]
    var = torch.var(g, dim=-2, keepdim=True).expand(-1, -1, S, -1)      # unbiased, nerf.py:82
    avg = torch.mean(g, dim=-2, keepdim=True).expand(-1, -1, S, -1)
    h = F.relu(_lin(sd, p + ".global_fc.0", torch.cat([g, var, avg], dim=-1)))
    w = F.softmax(F.relu(_lin(sd, p + ".agg_w_fc.0", h)), dim=-2)
    return F.relu(_lin(sd, p + ".fc.0", (h * w).sum(dim=-2)))


def nerf(sd, p, vox, f, viewdir_agg=True):
    """NeRF.forward, nerf.py:29-43.  vox (B,N,8), f (B,N,S,fc+4) -> (B,N,4) = rgb(3), sigma(1)."""
    S = f.shape[2]
    vif = torch.cat([vox, agg(sd, p + ".agg", f, viewdir_agg)], dim=-1)
    x = F.relu(_lin(sd, p + ".lr0.0", vif))
    sigma = F.softplus(_lin(sd, p + ".sigma.0", x))
    xx = torch.cat([x, vif], dim=-1)[:, :, None].expand(-1, -1, S, -1)
    c = F.relu(_lin(sd, p + ".color.2", F.relu(_lin(sd, p + ".color.0", torch.cat([xx, f], dim=-1)))))
    w = F.softmax(c, dim=-2)
    rgb = (f[..., -7:-4] * w).sum(dim=-2)
    return torch.cat([rgb, sigma], dim=-1)


# ------------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------------
def proj_mats(batch, src_scale, tar_scale):
    """get_proj_mats, utils.py:35-55 -> (B,S,3,4)."""
    B = batch["src_inps"].shape[0]
    k_s = batch["src_ixts"].clone()
    k_s[:, :, :2] *= src_scale
    p_s = k_s @ batch["src_exts"][:, :, :3]
    k_t = batch["tar_ixt"].clone()
    k_t[:, :2] *= tar_scale
    p_t = k_t @ batch["tar_ext"][:, :3]
    last = torch.zeros(B, 1, 4, dtype=p_t.dtype, device=p_t.device)
    last[:, :, 3] = 1
    p_t_inv = torch.inverse(torch.cat([p_t, last], dim=1))
    return p_s @ p_t_inv[:, None]


def homo_warp(src_feat, proj, depth_values):
    """homo_warp, utils.py:57-95.  src_feat (B,C,Hs,Ws), proj (B,3,4), depth (B,D,h,w)
    -> (B,C,D,h,w).  Samples the source at pixel coord xy (align_corners=True, zeros padding)."""
    B, D, h, w = depth_values.shape
    C, Hs, Ws = src_feat.shape[1:]
    dev = src_feat.device     # (device-agnostic: bench.py also runs this port on cuda:0 as the library-kernel baseline)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dev), torch.arange(w, dtype=torch.float32, device=dev), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=dev)], 0)[None].expand(B, -1, -1)  # (B,3,hw)
    pix = pix.repeat(1, 1, D)
    q = proj[:, :, :3] @ pix + proj[:, :, 3:] / depth_values.reshape(B, 1, D * h * w)
    xy = q[:, :2] / torch.clamp_min(q[:, 2:], 1e-6)
    gx = xy[:, 0] / ((Ws - 1) / 2) - 1
    gy = xy[:, 1] / ((Hs - 1) / 2) - 1
    grid = torch.stack([gx, gy], -1).view(B, D, h * w, 2)
    out = F.grid_sample(src_feat, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.view(B, C, D, h, w)


def depth_values_for_level(batch, cfg, level, D, depth, std, near_far):
    """get_depth_values, utils.py:98-151 -> (depth_values (B,D,h,w), near_far (B,2,h,w))."""
    c = cfg.enerf.cas_config
    B = batch["src_inps"].shape[0]
    H, W = batch["src_inps"].shape[-2:]
    h, w = int(H * c.volume_scale[level]), int(W * c.volume_scale[level])
    lin = torch.linspace(0.0, 1.0, steps=D, dtype=torch.float32, device=batch["src_inps"].device)
    if depth is None:
        nf = batch["near_far"]
        if c.depth_inv[level]:
            disp = 1.0 / nf[:, :1] + lin.view(1, -1).repeat(B, 1) * (1.0 / nf[:, 1:] - 1.0 / nf[:, :1])
            dv = 1.0 / disp
        else:
            dv = nf[:, :1] + (nf[:, 1:] - nf[:, :1]) * lin.view(1, -1).repeat(B, 1)
        dv = dv.view(B, D, 1, 1).repeat(1, 1, h, w)
    else:
        up = c.volume_scale[level] / c.volume_scale[level - 1]
        if up != 1.0:
            kw = dict(scale_factor=up, recompute_scale_factor=True, align_corners=True, mode="bilinear")
            depth = F.interpolate(depth[:, None], None, **kw)[:, 0]
            std = F.interpolate(std[:, None], None, **kw)[:, 0]
            near_far = F.interpolate(near_far, None, **kw)
        if not c.depth_inv[level - 1]:
            raise NotImplementedError("reference drops into ipdb here (utils.py:130)")
        lo = torch.minimum(depth + std, near_far[:, 0])   # utils.py:123-125 (masked clamp == min)
        hi = torch.maximum(depth - std, near_far[:, 1])   # utils.py:126-127
        nf = 1.0 / torch.stack([lo, hi], dim=-1)          # (B,h,w,2) metric [near, far]
        if c.depth_inv[level]:
            disp = 1.0 / nf[..., :1] + lin.view(1, 1, 1, -1) * (1.0 / nf[..., 1:] - 1.0 / nf[..., :1])
            dv = (1.0 / disp).permute(0, 3, 1, 2)
        else:
            dv = (nf[..., :1] + lin.view(1, 1, 1, -1) * (nf[..., 1:] - nf[..., :1])).permute(0, 3, 1, 2)
    out_nf = dv[:, [0, -1]]
    if c.depth_inv[level]:
        out_nf = 1.0 / torch.clamp_min(out_nf, 1e-6)
    return dv.contiguous(), out_nf


def feature_volume(feat, batch, cfg, level, depth, std, near_far, planes=None):
    """build_feature_volume, utils.py:322-349.  feat (B,S,C,hs,ws) -> variance (B,C,D,h,w).
    ``planes`` overrides cfg volume_planes[level] (the D argument of the reference function)."""
    c = cfg.enerf.cas_config
    S = feat.shape[1]
    dv, nf = depth_values_for_level(batch, cfg, level, planes or c.volume_planes[level], depth, std, near_far)
    pm = proj_mats(batch, c.im_feat_scale[level], c.volume_scale[level])
    s1, s2 = 0, 0
    for s in range(S):
        wv = homo_warp(feat[:, s], pm[:, s], dv)
        s1 = s1 + wv
        s2 = s2 + wv ** 2
    var = s2 / S - (s1 / S) ** 2          # utils.py:345, same association (Sum x^2 / S - (Sum x / S)^2)
    return var, dv, nf


def depth_regression(prob, dv, depth_inv):
    """depth_regression live part, utils.py:658-663 -> (depth, std) each (B,h,w)."""
    p = F.softmax(prob, 1)
    if depth_inv:
        dv = 1.0 / torch.clamp_min(dv, 1e-6)
    d = torch.sum(p * dv, 1)
    var = (p * (dv - d.unsqueeze(1)) ** 2).sum(1)
    return d, torch.clamp_min(var, 1e-10).sqrt()


def build_rays(depth, std, batch, cfg, near_far, level):
    """build_rays, utils.py:390-420 -> (B,N,12) = ray(8), ray_near_far(2), vol_near_far(2)."""
    c = cfg.enerf.cas_config
    up = c.render_scale[level] / c.volume_scale[level]
    if up != 1.0:
        depth = F.interpolate(depth[:, None], scale_factor=up, mode="bilinear", align_corners=True)[:, 0]
        std = F.interpolate(std[:, None], scale_factor=up, mode="bilinear", align_corners=True)[:, 0]
        near_far = F.interpolate(near_far, scale_factor=up, mode="bilinear", align_corners=True)
    if c.depth_inv[level]:
        a = torch.minimum(depth + std, near_far[:, 0])
        b = torch.maximum(depth - std, near_far[:, 1])
    else:
        a = torch.maximum(depth - std, near_far[:, 0])
        b = torch.minimum(depth + std, near_far[:, 1])
    rnf = torch.stack([a, b], dim=-1)                     # (B,Hr,Wr,2)
    vnf = near_far.permute(0, 2, 3, 1)
    rays = batch[f"rays_{level}"]
    uv = rays[:, :, 6:].long()
    rnf = torch.stack([rnf[i][uv[i][:, 1], uv[i][:, 0]] for i in range(len(rnf))])
    vnf = torch.stack([vnf[i][uv[i][:, 1], uv[i][:, 0]] for i in range(len(vnf))])
    return torch.cat([rays, rnf, vnf], dim=-1)


def sample_along_depth(rays, n_samples, depth_inv):
    """sample_along_depth, utils.py:422-441 -> xyz (B,N,Ns,3), uvd (B,N,Ns,3), z (B,N,Ns)."""
    o, d, uv = rays[..., :3], rays[..., 3:6], rays[..., 6:8]
    rn, rf, vn, vf = rays[..., 8:9], rays[..., 9:10], rays[..., 10:11], rays[..., 11:12]
    if n_samples == 1:
        z = rn + (rf - rn) * 0.5
    else:
        z = rn + (rf - rn) * torch.linspace(0.0, 1.0, n_samples, device=rays.device)[None, None]
    if depth_inv:
        xyz = o[..., None, :] + d[..., None, :] * (1 / torch.clamp_min(z[..., None], 1e-6))
        dn = (vn - z) / torch.clamp_min(vn - vf, 1e-6)
    else:
        xyz = o[..., None, :] + d[..., None, :] * z[..., None]
        dn = (z - vn) / torch.clamp_min(vf - vn, 1e-6)
    uvd = torch.cat([uv[..., None, :].repeat(1, 1, n_samples, 1), dn[..., None]], dim=-1)
    return xyz, uvd, z


def unpreprocess(src_inps, render_scale):
    """unpreprocess, utils.py:605-612."""
    img = src_inps * 0.5 + 0.5
    B, S, C, H, W = img.shape
    img = F.interpolate(img.reshape(B * S, C, H, W), scale_factor=render_scale, align_corners=True,
                        mode="bilinear", recompute_scale_factor=True)
    return img.reshape(B, S, C, int(H * render_scale), int(W * render_scale))


def vox_feat(uvd, vol):
    """get_vox_feat, utils.py:456-458.  uvd (B,P,3) in [0,1], vol (B,C,D,h,w) -> (B,P,C)."""
    return F.grid_sample(vol, uvd[:, None, None] * 2.0 - 1.0, align_corners=True)[:, :, 0, 0].permute(0, 2, 1)


def img_feat(xyz, img_feat_rgb, batch, render_scale):
    """get_img_feat, utils.py:689-722.  xyz (B,N,Ns,3), img_feat_rgb (B,S,C,H,W) -> (B,N*Ns,S,C+4)."""
    B, S, C, H, W = img_feat_rgb.shape
    pts = xyz.reshape(B, -1, 3)
    hom = torch.cat([pts, torch.ones_like(pts[..., :1])], dim=-1)
    tar_c = batch["tar_ext"].inverse()[:, :3, 3]
    out = []
    for s in range(S):
        cam = (hom @ batch["src_exts"][:, s].transpose(-1, -2))[..., :3]
        k = batch["src_ixts"][:, s].clone()
        k[:, :2] *= render_scale
        pix = cam @ k.transpose(-1, -2)
        g = pix[..., :2] / torch.clamp_min(pix[..., 2:], 1e-6)
        g = torch.stack([g[..., 0] / (W - 1), g[..., 1] / (H - 1)], -1) * 2.0 - 1.0
        f = F.grid_sample(img_feat_rgb[:, s], g[:, None], align_corners=True, mode="bilinear",
                          padding_mode="border").permute(0, 2, 3, 1)[:, 0]
        src_c = batch["src_exts"][:, s].inverse()[:, :3, 3]
        t = pts - tar_c[:, None]
        u = pts - src_c[:, None]
        t = t / (torch.norm(t, dim=-1, keepdim=True) + 1e-6)
        u = u / (torch.norm(u, dim=-1, keepdim=True) + 1e-6)
        r = t - u
        rn = torch.norm(r, dim=-1, keepdim=True)
        dot = torch.sum(t * u, dim=-1, keepdim=True)
        out.append(torch.cat([f, r / torch.clamp(rn, min=1e-6), dot], dim=-1))
    return torch.stack(out, -2)


def raw2outputs(raw, z, white_bkgd=False):
    """raw2outputs, utils.py:571-603 (alpha without dists; depth uses softmax(weights))."""
    alpha = 1.0 - torch.exp(-raw[..., 3])
    T = torch.cumprod(1.0 - alpha + 1e-10, dim=-1)[..., :-1]
    T = torch.cat([torch.ones_like(alpha[..., :1]), T], dim=-1)
    w = alpha * T
    rgb = torch.sum(w[..., None] * raw[..., :3], -2)
    w = F.softmax(w, dim=-1)
    depth = torch.sum(w * z, -1)
    if white_bkgd:
        rgb = rgb + (1.0 - torch.sum(w, -1)[..., None])
    return {"rgb": rgb, "depth": depth, "weights": w}


# ------------------------------------------------------------------------------------------------
# the forward
# ------------------------------------------------------------------------------------------------
def render_rays(sd, cfg, level, rays, batch, im_feat, vol, stash=None):
    """Network.render_rays, network.py:24-43 (one chunk)."""
    c = cfg.enerf.cas_config
    ns = c.num_samples[level]
    xyz, uvd, z = sample_along_depth(rays, ns, c.depth_inv[level])
    B = xyz.shape[0]
    rgbs = unpreprocess(batch["src_inps"], c.render_scale[level])
    upf = c.render_scale[level] / c.im_ibr_scale[level]
    if upf != 1.0:
        b, s, ch, hh, ww = im_feat.shape
        im_feat = F.interpolate(im_feat.reshape(b * s, ch, hh, ww), None, scale_factor=upf, align_corners=True,
                                mode="bilinear").view(b, s, ch, int(hh * upf), int(ww * upf))
    ifr = torch.cat([im_feat, rgbs], dim=2)
    Ho, Wo = batch["src_inps"].shape[-2:]
    Hr, Wr = int(Ho * c.render_scale[level]), int(Wo * c.render_scale[level])
    uvd = uvd.clone()
    uvd[..., 0] = uvd[..., 0] / (Wr - 1)
    uvd[..., 1] = uvd[..., 1] / (Hr - 1)
    vf = vox_feat(uvd.reshape(B, -1, 3), vol)
    ifd = img_feat(xyz, ifr, batch, c.render_scale[level])
    raw = nerf(sd, f"nerf_{level}", vf, ifd, cfg.enerf.viewdir_agg)
    raw = raw.reshape(B, -1, ns, raw.shape[-1])
    if stash is not None:
        stash.update(vox_feat=vf, img_feat_rgb_dir=ifd, raw=raw, z_vals=z, world_xyz=xyz)
    return raw2outputs(raw, z, cfg.enerf.white_bkgd)


def forward(sd, cfg, batch, intermediates=False, human=False):
    """Network.forward, network.py:76-113 (chunking by cfg.enerf.chunk_size as :45-55).
    ``human=True`` follows network_human.py:60-119 instead: identical except that, at the last level,
    only the rays inside ``batch['mask_at_box']`` are rendered (:90-92) and rgb is scattered back
    into a zero image (:102-106); depth / weights stay compact.
    Returns the reference's output dict; with ``intermediates`` also a dict of per-stage tensors."""
    c = cfg.enerf.cas_config
    B, S, _, H, W = batch["src_inps"].shape
    f2, f1, f0 = feature_net(sd, batch["src_inps"].reshape(B * S, 3, H, W))
    feats = {2: f0.reshape(B, S, -1, H, W), 1: f1.reshape(B, S, -1, H // 2, W // 2),
             0: f2.reshape(B, S, -1, H // 4, W // 4)}           # network.py:62-66
    ret, mid = {}, {"feat_level_0": feats[0], "feat_level_1": feats[1], "feat_level_2": feats[2]}
    depth = std = near_far = None
    for i in range(c.num):
        var, dv, near_far = feature_volume(feats[i], batch, cfg, i, depth, std, near_far)
        vol, prob = cost_reg(sd, f"cost_reg_{i}", var, deep=(i != 0))   # network.py:16-19
        depth, std = depth_regression(prob, dv, c.depth_inv[i])
        mid.update({f"variance_{i}": var, f"depth_values_{i}": dv, f"near_far_{i}": near_far,
                    f"feat_volume_{i}": vol, f"depth_prob_{i}": prob, f"depth_{i}": depth, f"std_{i}": std})
        if not c.render_if[i]:
            continue
        rays = build_rays(depth, std, batch, cfg, near_far, i)
        mid[f"rays12_{i}"] = rays
        masked = human and "mask_at_box" in batch and i == c.num - 1
        if masked:
            mask = batch["mask_at_box"].bool().reshape(1, -1)
            rays = rays[mask][None]
        chunk = int(cfg.enerf.chunk_size)
        parts, stash = [], ({} if intermediates else None)
        for j in range(0, rays.shape[1], chunk):
            parts.append(render_rays(sd, cfg, i, rays[:, j:j + chunk], batch, feats[c.render_im_feat_level[i]], vol,
                                     stash if j == 0 else None))
        out = {k: torch.cat([p[k] for p in parts], dim=1) for k in parts[0]}
        if stash:
            mid.update({f"{k}_{i}": v for k, v in stash.items()})
        if masked:
            rgb = torch.zeros(1, mask.shape[1], 3, device=out["rgb"].device)
            if mask.sum() > 1:
                rgb[mask] = out["rgb"][0]
            out["rgb"] = rgb
        out["depth_mvs"] = 1.0 / depth if c.depth_inv[i] else depth   # network.py:105-108
        out["std"] = std
        ret.update({f"{k}_level{i}": v for k, v in out.items()})
    return (ret, mid) if intermediates else ret
