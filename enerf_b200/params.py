"""Parameter tree of the drop-in ``Network``: same names, shapes, creation order and initialisers as
the reference, so ``state_dict()`` / ``load_state_dict(strict=True)`` round-trip with reference
checkpoints (SURVEY.md section 8b "Checkpoint keys"; /root/reference/lib/utils/net_utils.py:443).

The torch modules below are *parameter holders only*: the render path never calls their forward.
Kernels consume BN-folded, re-laid-out copies made by ``enerf_b200.packing``.

Tree (reference file:line that defines each group):
  feature_net.*            lib/networks/enerf/feature_net.py:5-22
  cost_reg_{i}.*           lib/networks/enerf/cost_reg_net.py:5-33 (CostRegNet), :52-73 (MinCostRegNet)
  nerf_{i}.{agg,lr0,sigma,color}   lib/networks/enerf/nerf.py:7-27, 46-72
"""
import torch.nn as nn

# (name, cin, cout, k, stride, pad) -- feature_net.py:7-15
FEATURE_CBR = [("conv0", [(3, 8, 3, 1, 1), (8, 8, 3, 1, 1)]),
               ("conv1", [(8, 16, 5, 2, 2), (16, 16, 3, 1, 1)]),
               ("conv2", [(16, 32, 5, 2, 2), (32, 32, 3, 1, 1)])]
# (name, cin, cout, k, pad) -- feature_net.py:17-22
FEATURE_PLAIN = [("toplayer", 32, 32, 1, 0), ("lat1", 16, 32, 1, 0), ("lat0", 8, 32, 1, 0),
                 ("smooth1", 32, 16, 3, 1), ("smooth0", 32, 8, 3, 1)]


def cost_reg_layers(in_ch, deep):
    """Layer spec of (Min)CostRegNet: list of (name, kind, cin, cout, stride)."""
    spec = [("conv0", "cbr", in_ch, 8, 1), ("conv1", "cbr", 8, 16, 2), ("conv2", "cbr", 16, 16, 1),
            ("conv3", "cbr", 16, 32, 2), ("conv4", "cbr", 32, 32, 1)]
    if deep:
        spec += [("conv5", "cbr", 32, 64, 2), ("conv6", "cbr", 64, 64, 1), ("conv7", "deconv", 64, 32, 2)]
    spec += [("conv9", "deconv", 32, 16, 2), ("conv11", "deconv", 16, 8, 2),
             ("depth_conv", "plain", 8, 1, 1), ("feat_conv", "plain", 8, 8, 1)]
    return spec


class ConvBN(nn.Module):
    """Holder for conv (no bias) + batch-norm; keys ``conv.weight`` / ``bn.*`` (utils.py:10-33)."""

    def __init__(self, dims, cin, cout, k, stride, pad):
        super().__init__()
        conv = nn.Conv2d if dims == 2 else nn.Conv3d
        bn = nn.BatchNorm2d if dims == 2 else nn.BatchNorm3d
        self.conv = conv(cin, cout, k, stride=stride, padding=pad, bias=False)
        self.bn = bn(cout)


class FeatureParams(nn.Module):
    def __init__(self):
        super().__init__()
        for name, pair in FEATURE_CBR:
            setattr(self, name, nn.Sequential(*[ConvBN(2, *p) for p in pair]))
        for name, cin, cout, k, pad in FEATURE_PLAIN:
            setattr(self, name, nn.Conv2d(cin, cout, k, padding=pad))


class CostRegParams(nn.Module):
    def __init__(self, in_ch, deep):
        super().__init__()
        self.in_ch, self.deep = in_ch, deep
        for name, kind, cin, cout, stride in cost_reg_layers(in_ch, deep):
            if kind == "cbr":
                m = ConvBN(3, cin, cout, 3, stride, 1)
            elif kind == "deconv":
                m = nn.Sequential(nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False),
                                  nn.BatchNorm3d(cout))
            else:
                m = nn.Sequential(nn.Conv3d(cin, cout, 3, padding=1, bias=False))
            setattr(self, name, m)


def _kaiming(m):
    # nerf.py:130-134
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


def _fc(cin, cout):
    return nn.Sequential(nn.Linear(cin, cout))


class AggParams(nn.Module):
    def __init__(self, feat_ch, viewdir_agg):
        super().__init__()
        self.feat_ch = feat_ch
        if viewdir_agg:
            self.view_fc = _fc(4, feat_ch)
            self.view_fc.apply(_kaiming)
        self.global_fc = _fc(feat_ch * 3, 32)
        self.agg_w_fc = _fc(32, 1)
        self.fc = _fc(32, 16)
        for m in (self.global_fc, self.agg_w_fc, self.fc):
            m.apply(_kaiming)


class NerfParams(nn.Module):
    def __init__(self, feat_ch, viewdir_agg, hid=64, vox_ch=8):
        """vox_ch=8: nerf.NeRF (nerf.py:7-27); vox_ch=0: nerf_.NeRF, which drops the voxel feature
        (nerf_.py:13,20: lr0 Linear(16,hid), color.0 Linear(hid+16+feat_ch+4,hid))."""
        super().__init__()
        self.feat_ch, self.vox_ch = feat_ch, vox_ch
        self.agg = AggParams(feat_ch, viewdir_agg)
        self.lr0 = _fc(vox_ch + 16, hid)
        self.sigma = _fc(hid, 1)
        # indices 0 and 2 carry the Linear layers (nerf.py:21-25: Linear, ReLU, Linear, ReLU)
        self.color = nn.Sequential(nn.Linear(hid + vox_ch + 16 + feat_ch + 4, hid), nn.Identity(), nn.Linear(hid, 1))
        for m in (self.lr0, self.sigma, self.color):
            m.apply(_kaiming)
