"""Weight packing for the CUDA kernels: fold eval-mode BatchNorm into the preceding convolution and
re-lay every tensor into the order the kernels read (include/enerf_b200.h documents each layout).

BN folding (eval mode, eps 1e-5; /root/reference/lib/networks/enerf/utils.py:10-33):
    s = gamma / sqrt(running_var + eps);   w' = w * s[cout];   b' = beta - running_mean * s
ConvTranspose3d weights are (Cin, Cout, k, k, k), so the scale runs along dim 1
(cost_reg_net.py:19-33).  Folding is done in float64 and rounded once to float32.
"""
import torch

from .params import FEATURE_CBR, FEATURE_PLAIN, cost_reg_layers

BN_EPS = 1e-5


def _fold(w, sd, bn, out_dim):
    scale = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + BN_EPS)
    shape = [1] * w.dim()
    shape[out_dim] = -1
    return w.double() * scale.view(shape), sd[bn + ".bias"].double() - sd[bn + ".running_mean"].double() * scale


def _taps_cin_cout(w):
    """(Cout, Cin, k...) -> [tap][cin][cout]"""
    cout, cin = w.shape[:2]
    perm = list(range(2, w.dim())) + [1, 0]
    return w.permute(perm).reshape(-1, cin, cout)


def _dev(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


# layers enerf_feature_net runs on tcgen05 when tensor_cores != 0 (every layer with cin % 8 == 0; all but conv0.0)
TC_FEATURE_LAYERS = ("conv0.1", "conv1.0", "conv1.1", "conv2.0", "conv2.1", "toplayer", "smooth1", "smooth0")


# Which layers carry their three kx taps in the MMA's N dimension; mirrors csrc/tc_conv.cu::tc_fold_rule (capi.tc_conv_fold_rule
# sets both): 0 = stride-1 3x3x3 layers with 8 output channels + the single-channel depth head, 1 = + the feat/prob head,
# 2 (shipped) = + stride-1 3x3 2-D layers with 8 output and >= 16 input channels (FeatureNet smooth0).
FOLD_RULE = 2


def tc_fold_kx(KD, KH, stride, cout, single=False, head=False, cin=32):
    """The folding rule of csrc/tc_conv.cu::tc_fold_rule for a convolution (KD x KH x KH, `stride`, `cin` -> `cout` channels;
    single: the depth-only head, head: the feat + prob head)."""
    if stride != 1 or KH != 3:
        return False
    if KD == 3:
        return bool(single or (head and FOLD_RULE >= 1) or (not head and cout == 8))
    return KD == 1 and not single and not head and cout == 8 and cin >= 16 and FOLD_RULE >= 2


def pack_feature_net(sd, device, p="feature_net", tensor_cores=False):
    """22 tensors: {w,b} for conv0.0 conv0.1 conv1.0 conv1.1 conv2.0 conv2.1 toplayer lat1 lat0 smooth1 smooth0.
    tensor_cores: the layers in TC_FEATURE_LAYERS get the tcgen05 stage layout (pack_tc_conv)."""
    def lay(name, w_taps):
        if tensor_cores and name in TC_FEATURE_LAYERS:
            fold = w_taps.shape[0] == 9 and tc_fold_kx(1, 3, 1, w_taps.shape[2], cin=w_taps.shape[1])     # FeatureNet's 3x3 layers are all stride 1
            return pack_tc_conv(w_taps, fold_kx=fold).to(device)
        return _dev(w_taps, device)

    out = []
    for name, pair in FEATURE_CBR:
        for j in range(len(pair)):
            q = f"{p}.{name}.{j}"
            w, b = _fold(sd[q + ".conv.weight"], sd, q + ".bn", 0)
            out += [lay(f"{name}.{j}", _taps_cin_cout(w)), _dev(b, device)]
    for name, *_ in FEATURE_PLAIN:
        out += [lay(name, _taps_cin_cout(sd[f"{p}.{name}.weight"].double())), _dev(sd[f"{p}.{name}.bias"], device)]
    return out


def pack_cost_reg(sd, p, in_ch, deep, device, with_feat, tensor_cores=False):
    """{w,b} per layer (conv0..conv11) + one head tensor ([27][8][9] feat+depth, or [27][8][1] depth only).
    tensor_cores: stride-1 and transposed layers + head in the tcgen05 stage layout."""
    out = []
    head = {}
    for name, kind, cin, cout, stride in cost_reg_layers(in_ch, deep):
        q = f"{p}.{name}"
        if kind == "cbr":
            w, b = _fold(sd[q + ".conv.weight"], sd, q + ".bn", 0)
            wt = _taps_cin_cout(w)
            use_tc = tensor_cores
            out += [pack_tc_conv(wt, fold_kx=tc_fold_kx(3, 3, stride, cout)).to(device) if use_tc else _dev(wt, device), _dev(b, device)]
        elif kind == "deconv":
            w, b = _fold(sd[q + ".0.weight"], sd, q + ".1", 1)      # (Cin,Cout,kz,ky,kx)
            wt = w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout)
            out += [pack_tc_deconv(wt).to(device) if tensor_cores else _dev(wt, device), _dev(b, device)]
        else:
            head[name] = _taps_cin_cout(sd[q + ".0.weight"].double())   # [27][8][cout]
    hw = torch.cat([head["feat_conv"], head["depth_conv"]], dim=2) if with_feat else head["depth_conv"]
    out.append(pack_tc_conv(hw, fold_kx=tc_fold_kx(3, 3, 1, hw.shape[-1], single=not with_feat, head=with_feat)).to(device) if tensor_cores
               else _dev(hw, device))
    return out


def pack_nerf(sd, p, feat_ch, viewdir_agg, device):
    """16 tensors, every Linear transposed to [in][out] (see enerf_render_rays in enerf_b200.h)."""
    def lin(q):
        return [_dev(sd[q + ".weight"].t(), device), _dev(sd[q + ".bias"], device)]

    out = []
    if viewdir_agg:
        out += lin(f"{p}.agg.view_fc.0")
    else:
        out += [None, None]
    for q in ("agg.global_fc.0", "agg.agg_w_fc.0", "agg.fc.0", "lr0.0", "sigma.0", "color.0", "color.2"):
        out += lin(f"{p}.{q}")
    return out


def _novox_as_vox(sd, p):
    """nerf_.NeRF (no voxel feature, nerf_.py:29-43) embedded in nerf.NeRF's shapes: zero columns
    where nerf.NeRF reads the 8 voxel channels (lr0 inputs 0..7, color.0 inputs 64..71), so the ray
    kernels compute exactly the nerf_ network (adding 0 * x is exact)."""
    sub = {k: v for k, v in sd.items() if k.startswith(p + ".")}
    lr0, c0 = sub[f"{p}.lr0.0.weight"], sub[f"{p}.color.0.weight"]      # (64,16), (64, 64+16+fc+4)
    z8 = torch.zeros((64, 8), device=lr0.device, dtype=lr0.dtype)
    sub[f"{p}.lr0.0.weight"] = torch.cat([z8, lr0], dim=1)
    sub[f"{p}.color.0.weight"] = torch.cat([c0[:, :64], z8, c0[:, 64:]], dim=1)
    return sub


def pack_nerf_novox(sd, p, feat_ch, viewdir_agg, device):
    """nerf_.NeRF in the 16-tensor layout of ``pack_nerf`` (for enerf_render_rays_raw)."""
    return pack_nerf(_novox_as_vox(sd, p), p, feat_ch, viewdir_agg, device)


def pack_nerf_tc_novox(sd, p, feat_ch, viewdir_agg, device):
    """nerf_.NeRF in the blob layout of ``pack_nerf_tc`` (for enerf_render_rays_raw_tc)."""
    return pack_nerf_tc(_novox_as_vox(sd, p), p, feat_ch, viewdir_agg, device)


def tf32_round(x):
    """cvt.rna.tf32.f32 on the host: round to nearest (ties away from zero) to a 10-bit mantissa."""
    b = x.to(torch.float32).contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def _b_chunks(wt, K, N):
    """W^T [k][n] (zero padded to K x N) -> tcgen05 K-major no-swizzle chunk layout [K/4][N][4], TF32."""
    full = torch.zeros(K, N, dtype=torch.float32)
    full[: wt.shape[0], : wt.shape[1]] = wt
    return tf32_round(full.view(K // 4, 4, N).permute(0, 2, 1).contiguous()).reshape(-1)


def pack_nerf_tc(sd, p, feat_ch, viewdir_agg, device):
    """One 10,392-float blob for enerf_render_rays_tc (layout: struct TcW in csrc/render_rays_tc.cu)."""
    fc = feat_ch
    assert fc == 11, "tensor-core ray kernel is built for feat_ch 8 (+3 rgb)"
    cpu = {k: v.detach().float().cpu() for k, v in sd.items() if k.startswith(p + ".")}
    gw = cpu[f"{p}.agg.global_fc.0.weight"]            # (32, 3*fc): [f | var | mean]
    cw = cpu[f"{p}.color.0.weight"]                    # (64, 88 + fc + 4)
    shared = torch.zeros(24, 32)
    shared[0:fc] = gw[:, fc:2 * fc].t()
    shared[12:12 + fc] = gw[:, 2 * fc:3 * fc].t()
    # the biases of global_fc and color.0 ride in the GEMMs: the kernel stores a constant 1 in the
    # first padding column of the per-view A operands (col 11 of [g_s|pad], col 15 of [f_s|dir_s|pad])
    gv = torch.zeros(16, 32)
    gv[:fc] = gw[:, :fc].t()
    gv[fc] = cpu[f"{p}.agg.global_fc.0.bias"]
    cv = torch.zeros(16, 64)
    cv[:fc + 4] = cw[:, 88:88 + fc + 4].t()
    cv[fc + 4] = cpu[f"{p}.color.0.bias"]
    parts = [
        _b_chunks(gv, 16, 32),
        _b_chunks(shared, 24, 32),
        _b_chunks(cpu[f"{p}.agg.fc.0.weight"].t(), 32, 16),
        _b_chunks(cpu[f"{p}.lr0.0.weight"].t(), 24, 64),
        _b_chunks(cw[:, :88].t(), 88, 64),
        _b_chunks(cv, 16, 64),
    ]
    view_w, view_b = torch.zeros(4, 12), torch.zeros(12)
    if viewdir_agg:
        view_w[:, :fc] = cpu[f"{p}.agg.view_fc.0.weight"].t()
        view_b[:fc] = cpu[f"{p}.agg.view_fc.0.bias"]

    def pad4(t):
        out = torch.zeros(4)
        out[: t.numel()] = t.reshape(-1)
        return out

    parts += [view_w.reshape(-1), view_b, cpu[f"{p}.agg.global_fc.0.bias"], cpu[f"{p}.agg.agg_w_fc.0.weight"].reshape(-1),
              pad4(cpu[f"{p}.agg.agg_w_fc.0.bias"]), cpu[f"{p}.agg.fc.0.bias"], cpu[f"{p}.lr0.0.bias"],
              cpu[f"{p}.sigma.0.weight"].reshape(-1), pad4(cpu[f"{p}.sigma.0.bias"]), cpu[f"{p}.color.0.bias"],
              cpu[f"{p}.color.2.weight"].reshape(-1), pad4(cpu[f"{p}.color.2.bias"])]
    blob = torch.cat([t.reshape(-1).float() for t in parts])
    assert blob.numel() == 10392, blob.numel()
    return blob.to(device).contiguous()


def pack_tc_conv(w_taps, n_pad=None, fold_kx=False):
    """[tap][cin][cout] (BN-folded, fp32) -> tcgen05 stage layout [cin/8][tap][2][N][4], TF32 (RNA).
    fold_kx (stride-1 3x3 / 3x3x3 layers, the rule tc_conv_launch applies): the three kx taps move into
    the N dimension, [KD*KH][cin][kx*cout + co], N = 3*cout padded to 16 (csrc/tc_conv.cu, FOLD)."""
    if fold_kx:
        taps, cin, cout = w_taps.shape
        assert taps in (9, 27)
        w_taps = w_taps.reshape(taps // 3, 3, cin, cout).permute(0, 2, 1, 3).reshape(taps // 3, cin, 3 * cout)
    taps, cin, cout = w_taps.shape
    N = n_pad or (cout + 15) // 16 * 16
    full = torch.zeros(taps, cin, N, dtype=torch.float32)
    full[:, :, :cout] = w_taps.float().cpu()
    # cin = st*8 + j*4 + i
    return tf32_round(full.view(taps, cin // 8, 2, 4, N).permute(1, 0, 2, 4, 3).contiguous()).reshape(-1)


def pack_tc_deconv(w_taps):
    """ConvTranspose3d(k3,s2,p1,op1) [27][cin][cout] -> sub-pixel conv: 8 input offsets d in {0,1}^3,
    N = 8 parities x cout.  Per dimension: offset 0 feeds parity 0 through k=1 and parity 1 through
    k=2; offset 1 feeds parity 1 through k=0 (o = 2i - 1 + k)."""
    taps, cin, cout = w_taps.shape
    assert taps == 27
    w = w_taps.float().cpu().view(3, 3, 3, cin, cout)
    kmap = {(0, 0): 1, (0, 1): 2, (1, 1): 0}
    full = torch.zeros(8, cin, 8, cout, dtype=torch.float32)      # [offset][cin][parity][cout]
    for d in range(8):
        dz, dy, dx = (d >> 2) & 1, (d >> 1) & 1, d & 1
        for e in range(8):
            ez, ey, ex = (e >> 2) & 1, (e >> 1) & 1, e & 1
            if (dz, ez) in kmap and (dy, ey) in kmap and (dx, ex) in kmap:
                full[d, :, e, :] = w[kmap[(dz, ez)], kmap[(dy, ey)], kmap[(dx, ex)]]
    return pack_tc_conv(full.view(8, cin, 8 * cout))
