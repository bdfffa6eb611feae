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


def pack_feature_net(sd, device, p="feature_net"):
    """22 tensors: {w,b} for conv0.0 conv0.1 conv1.0 conv1.1 conv2.0 conv2.1 toplayer lat1 lat0 smooth1 smooth0."""
    out = []
    for name, pair in FEATURE_CBR:
        for j in range(len(pair)):
            q = f"{p}.{name}.{j}"
            w, b = _fold(sd[q + ".conv.weight"], sd, q + ".bn", 0)
            out += [_dev(_taps_cin_cout(w), device), _dev(b, device)]
    for name, *_ in FEATURE_PLAIN:
        out += [_dev(_taps_cin_cout(sd[f"{p}.{name}.weight"].double()), device), _dev(sd[f"{p}.{name}.bias"], device)]
    return out


def pack_cost_reg(sd, p, in_ch, deep, device, with_feat):
    """{w,b} per layer (conv0..conv11) + one head tensor ([27][8][9] feat+depth, or [27][8][1] depth only)."""
    out = []
    head = {}
    for name, kind, cin, cout, stride in cost_reg_layers(in_ch, deep):
        q = f"{p}.{name}"
        if kind == "cbr":
            w, b = _fold(sd[q + ".conv.weight"], sd, q + ".bn", 0)
            out += [_dev(_taps_cin_cout(w), device), _dev(b, device)]
        elif kind == "deconv":
            w, b = _fold(sd[q + ".0.weight"], sd, q + ".1", 1)      # (Cin,Cout,kz,ky,kx)
            out += [_dev(w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout), device), _dev(b, device)]
        else:
            head[name] = _taps_cin_cout(sd[q + ".0.weight"].double())   # [27][8][cout]
    if with_feat:
        out.append(_dev(torch.cat([head["feat_conv"], head["depth_conv"]], dim=2), device))
    else:
        out.append(_dev(head["depth_conv"], device))
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
