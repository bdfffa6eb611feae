"""Per-stage parity harness: runs every C-ABI entry point on the B200 fed with the ORACLE's
inputs for that stage, and the full forward end to end, and reports max-abs errors.  Used by
tests/test_parity_gpu.py and as a CLI (`python tests/stage_harness.py <golden-case|HxWxS>`)
that dumps a JSON report into gpurun_out/."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from enerf_b200 import capi, packing, synthetic  # noqa: E402
from enerf_b200 import config as bcfg  # noqa: E402
from enerf_b200.config import snapshot  # noqa: E402
from oracle import enerf_oracle as O  # noqa: E402


def nhwc(x):  # (S,C,H,W) -> (S,H,W,C) cuda
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def ndhwc(x):  # (C,D,H,W) -> (D,H,W,C) cuda
    return x.permute(1, 2, 3, 0).contiguous().cuda()


def err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return {"max_abs": (a - b).abs().max().item(), "ref_max": b.abs().max().item(),
            "nan": bool(torch.isnan(a).any().item())}


def make_case(H, W, S, cfg, seed=2):
    from enerf_b200.network import Network
    bcfg.set_cfg(cfg)
    torch.manual_seed(0)
    net = Network().eval()
    synthetic.randomize_bn_(net, seed=1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    batch = synthetic.make_batch(H, W, S, cfg, seed=seed)
    return sd, batch


def stage_report(sd, cfg, batch, oracle_out=None, oracle_mid=None, precision="fp32", human=False):
    """Returns {stage: err-dict}.  Needs a CUDA device."""
    from enerf_b200.network import Network
    bcfg.set_cfg(cfg)
    if oracle_out is None:
        with torch.no_grad():
            oracle_out, oracle_mid = O.forward(sd, cfg, batch, intermediates=True, human=human)
    mid = oracle_mid
    levels = snapshot(cfg)
    dev = torch.device("cuda")
    f32 = dict(device=dev, dtype=torch.float32)
    rep = {}
    S, _, H, W = batch["src_inps"][0].shape
    gb = {k: v.cuda() for k, v in batch.items()}

    # camera
    cam = torch.zeros(capi.CAM_FLOATS, **f32)
    scales = [(lv.im_feat_scale, lv.volume_scale, lv.render_scale) for lv in levels]
    capi.camera_setup(gb["src_exts"][0].contiguous(), gb["src_ixts"][0].contiguous(), gb["tar_ext"][0].contiguous(),
                      gb["tar_ixt"][0].contiguous(), gb["near_far"][0].contiguous(), scales, cam)
    homo = cam[: capi.MAX_LEVELS * capi.MAX_VIEWS * 12].view(capi.MAX_LEVELS, capi.MAX_VIEWS, 3, 4)
    for i, lv in enumerate(levels):
        pm = O.proj_mats(batch, lv.im_feat_scale, lv.volume_scale)[0]
        rep[f"camera.homo{i}"] = err(homo[i, :S] / pm.abs().max(), pm / pm.abs().max())

    # feature net
    pk_feat = packing.pack_feature_net(sd, dev)
    f0 = torch.empty((S, H // 4, W // 4, 32), **f32)
    f1 = torch.empty((S, H // 2, W // 2, 16), **f32)
    f2 = torch.empty((S, H, W, 8), **f32)
    ws = torch.empty(capi.feature_net_workspace_bytes(S, H, W) // 4, **f32)
    capi.feature_net(pk_feat, gb["src_inps"][0].contiguous(), f0, f1, f2, ws)
    for i, t in enumerate((f0, f1, f2)):
        rep[f"feature_net.level_{i}"] = err(t.permute(0, 3, 1, 2), mid[f"feat_level_{i}"][0])
    # same stage with the stride-1 layers on the tensor cores (TF32)
    pk_feat_tc = packing.pack_feature_net(sd, dev, tensor_cores=True)
    g0, g1, g2 = torch.full_like(f0, float("nan")), torch.full_like(f1, float("nan")), torch.full_like(f2, float("nan"))
    capi.feature_net(pk_feat_tc, gb["src_inps"][0].contiguous(), g0, g1, g2, ws, tensor_cores=True)
    for i, t in enumerate((g0, g1, g2)):
        rep[f"feature_net_tc.level_{i}"] = err(t.permute(0, 3, 1, 2), mid[f"feat_level_{i}"][0])

    prev = None
    for i, lv in enumerate(levels):
        h, w, D = int(H * lv.volume_scale), int(W * lv.volume_scale), lv.planes
        deep = i != 0
        ends = torch.empty((2, h, w), **f32)
        nf = torch.empty((2, h, w), **f32)
        if prev is None:
            capi.depth_hypotheses(cam, None, None, None, h, w, D, lv.depth_inv, ends, nf)
        else:
            capi.depth_hypotheses(cam, prev[0], prev[1], prev[2], h, w, D, lv.depth_inv, ends, nf)
        rep[f"depth_hypotheses.near_far_{i}"] = err(nf, mid[f"near_far_{i}"][0])
        rep[f"depth_hypotheses.ends_{i}"] = err(ends, mid[f"depth_values_{i}"][0][[0, -1]])
        # cost volume fed with oracle features
        feat = nhwc(mid[f"feat_level_{i}"][0])
        C = feat.shape[-1]
        var = torch.empty((D, h, w, C), **f32)
        capi.cost_volume(cam, i, feat, ends, D, h, w, lv.depth_inv, var)
        rep[f"cost_volume.variance_{i}"] = err(var.permute(3, 0, 1, 2), mid[f"variance_{i}"][0])
        # cost reg fed with oracle variance
        pk_reg = packing.pack_cost_reg(sd, f"cost_reg_{i}", int(32 * 2 ** (-i)), deep, dev, True)
        vol = torch.empty((D, h, w, 8), **f32)
        prob = torch.empty((D, h, w), **f32)
        rws = torch.empty(capi.cost_reg_workspace_bytes(deep, D, h, w) // 4, **f32)
        capi.cost_reg(pk_reg, deep, ndhwc(mid[f"variance_{i}"][0]), vol, prob, rws)
        rep[f"cost_reg.feat_volume_{i}"] = err(vol.permute(3, 0, 1, 2), mid[f"feat_volume_{i}"][0])
        rep[f"cost_reg.depth_prob_{i}"] = err(prob, mid[f"depth_prob_{i}"][0])
        pk_reg1 = packing.pack_cost_reg(sd, f"cost_reg_{i}", int(32 * 2 ** (-i)), deep, dev, False)
        prob1 = torch.empty((D, h, w), **f32)
        capi.cost_reg(pk_reg1, deep, ndhwc(mid[f"variance_{i}"][0]), None, prob1, rws)
        rep[f"cost_reg.depth_prob_only_{i}"] = err(prob1, mid[f"depth_prob_{i}"][0])
        for with_feat in (True, False):
            pk_tc = packing.pack_cost_reg(sd, f"cost_reg_{i}", int(32 * 2 ** (-i)), deep, dev, with_feat, tensor_cores=True)
            vol_t = torch.full((D, h, w, 8), float("nan"), **f32) if with_feat else None
            prob_t = torch.full((D, h, w), float("nan"), **f32)
            capi.cost_reg(pk_tc, deep, ndhwc(mid[f"variance_{i}"][0]), vol_t, prob_t, rws, tensor_cores=True)
            if with_feat:
                rep[f"cost_reg_tc.feat_volume_{i}"] = err(vol_t.permute(3, 0, 1, 2), mid[f"feat_volume_{i}"][0])
                rep[f"cost_reg_tc.depth_prob_{i}"] = err(prob_t, mid[f"depth_prob_{i}"][0])
            else:
                rep[f"cost_reg_tc.depth_prob_only_{i}"] = err(prob_t, mid[f"depth_prob_{i}"][0])
        # depth regression fed with oracle prob (and the oracle-equivalent ends computed above)
        depth = torch.empty((h, w), **f32)
        std = torch.empty((h, w), **f32)
        mvs = torch.empty((h, w), **f32)
        capi.depth_regress(mid[f"depth_prob_{i}"][0].contiguous().cuda(), ends, lv.depth_inv, depth, std, mvs)
        rep[f"depth_regress.depth_{i}"] = err(depth, mid[f"depth_{i}"][0])
        rep[f"depth_regress.std_{i}"] = err(std, mid[f"std_{i}"][0])
        o_depth, o_std, o_nf = (mid[f"depth_{i}"][0].contiguous().cuda(), mid[f"std_{i}"][0].contiguous().cuda(),
                                mid[f"near_far_{i}"][0].contiguous().cuda())
        prev = (o_depth, o_std, o_nf)
        if not lv.render_if or (human and i == len(levels) - 1):
            continue          # (the masked level is covered end to end below)
        # fused ray stage fed with oracle depth/std/near_far/volume/features
        Hr, Wr = int(H * lv.render_scale), int(W * lv.render_scale)
        imf = nhwc(mid[f"feat_level_{lv.im_feat_level}"][0])
        img = torch.empty((S, Hr, Wr, lv.feat_ch + 4), **f32)
        capi.pack_img_feat(imf, gb["src_inps"][0].contiguous(), img)
        pk_nerf = packing.pack_nerf(sd, f"nerf_{i}", lv.feat_ch + 3, bool(cfg.enerf.viewdir_agg), dev)
        rays = gb[f"rays_{i}"][0].contiguous()
        N = rays.shape[0]
        rgb = torch.empty((N, 3), **f32)
        dmap = torch.empty((N,), **f32)
        wts = torch.empty((N, lv.num_samples), **f32)
        capi.render_rays(cam, i, pk_nerf, rays, o_depth, o_std, o_nf, ndhwc(mid[f"feat_volume_{i}"][0]), img, lv.feat_ch,
                         lv.num_samples, lv.depth_inv, bool(cfg.enerf.white_bkgd), bool(cfg.enerf.viewdir_agg), rgb, dmap, wts)
        rep[f"render_rays.rgb_{i}"] = err(rgb, oracle_out[f"rgb_level{i}"][0])
        rep[f"render_rays.depth_{i}"] = err(dmap, oracle_out[f"depth_level{i}"][0])
        rep[f"render_rays.weights_{i}"] = err(wts, oracle_out[f"weights_level{i}"][0])
        if capi.tc_ray_kernel_supports(lv.feat_ch, S, lv.num_samples):
            blob = packing.pack_nerf_tc(sd, f"nerf_{i}", lv.feat_ch + 3, bool(cfg.enerf.viewdir_agg), dev)
            rgb2, d2, w2 = torch.full_like(rgb, float("nan")), torch.full_like(dmap, float("nan")), torch.full_like(wts, float("nan"))
            capi.render_rays_tc(cam, i, blob, rays, o_depth, o_std, o_nf, ndhwc(mid[f"feat_volume_{i}"][0]), img, lv.feat_ch,
                                lv.num_samples, lv.depth_inv, bool(cfg.enerf.white_bkgd), bool(cfg.enerf.viewdir_agg), rgb2, d2, w2)
            rep[f"render_rays_tc.rgb_{i}"] = err(rgb2, oracle_out[f"rgb_level{i}"][0])
            rep[f"render_rays_tc.depth_{i}"] = err(d2, oracle_out[f"depth_level{i}"][0])
            rep[f"render_rays_tc.weights_{i}"] = err(w2, oracle_out[f"weights_level{i}"][0])
            if S <= 3:   # 2-3 views: the warp-specialised kernel (render_rays_ws.cu, opt-in) must agree as well
                capi.render_rays_tc_select(2)
                try:
                    rgb3, d3, w3 = torch.full_like(rgb, float("nan")), torch.full_like(dmap, float("nan")), torch.full_like(wts, float("nan"))
                    capi.render_rays_tc(cam, i, blob, rays, o_depth, o_std, o_nf, ndhwc(mid[f"feat_volume_{i}"][0]), img, lv.feat_ch,
                                        lv.num_samples, lv.depth_inv, bool(cfg.enerf.white_bkgd), bool(cfg.enerf.viewdir_agg), rgb3, d3, w3)
                    torch.cuda.synchronize()
                finally:
                    capi.render_rays_tc_select(0)
                rep[f"render_rays_tc.ws_rgb_{i}"] = err(rgb3, oracle_out[f"rgb_level{i}"][0])
                rep[f"render_rays_tc.ws_depth_{i}"] = err(d3, oracle_out[f"depth_level{i}"][0])
                rep[f"render_rays_tc.ws_weights_{i}"] = err(w3, oracle_out[f"weights_level{i}"][0])

    # end to end through the drop-in Network
    if human:
        from enerf_b200.network_human import Network as HumanNetwork
        net = HumanNetwork()
    else:
        net = Network()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    net.precision = precision
    with torch.no_grad():
        out = net(gb)
    torch.cuda.synchronize()
    assert set(out) == set(oracle_out), (sorted(out), sorted(oracle_out))
    for k in oracle_out:
        assert out[k].shape == oracle_out[k].shape, (k, out[k].shape, oracle_out[k].shape)
        rep[f"e2e.{k}"] = err(out[k], oracle_out[k])
    for i, lv in enumerate(levels):
        if lv.render_if:
            tgt = torch.rand(oracle_out[f"rgb_level{i}"].shape, generator=torch.Generator().manual_seed(5))
            p_ours = synthetic.psnr(out[f"rgb_level{i}"].cpu(), tgt)
            p_ref = synthetic.psnr(oracle_out[f"rgb_level{i}"], tgt)
            rep[f"e2e.psnr_level{i}"] = {"psnr_ours_vs_ref": synthetic.psnr(out[f"rgb_level{i}"].cpu(), oracle_out[f"rgb_level{i}"]),
                                         "delta_psnr": p_ours - p_ref}
    return rep, out


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _helpers import load_golden

    arg = sys.argv[1] if len(sys.argv) > 1 else "c2_small_cascade"
    if "x" in arg:
        H, W, S = (int(v) for v in arg.split("x"))
        cfg = bcfg.make_cfg(volume_planes=[48, 8], render_if=[False, True])
        sd, batch = make_case(H, W, S, cfg)
        rep, _ = stage_report(sd, cfg, batch)
    else:
        fx = load_golden(arg)
        rep, _ = stage_report(fx["state_dict"], fx["cfg"], fx["batch"], human=fx.get("human", False))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"stage_report_{arg}.json"), "w") as f:
        json.dump(rep, f, indent=1)
    for k, v in rep.items():
        print(f"{k:40s} {v}")
