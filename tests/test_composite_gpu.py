"""GPU parity tests of the layered ("composite") path: the windowed
stages and the per-pixel merge against the CPU oracle (oracle/enerf_oracle_composite.py), and the
drop-in enerf_b200.network_composite.Network end to end against what the unmodified reference
produced (tests/golden/c5_*.pt).

Tolerances: FP32-pipe build ("fp32"): 5e-4 abs on rgb / weights / net_output, 2e-3 * max|ref| on
depth and z; TF32 conv stacks ("tf32", default): 2e-3 / 4e-3 (same bounds as the single-layer path).
"""
import pytest
import torch

from _helpers import COMPOSITE_CASES, load_golden
from test_oracle_golden import check_composite_outputs

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _to_dev(batch):
    return {k: (v.cuda() if torch.is_tensor(v) and k != "bbox" else v) for k, v in batch.items()}


def _make_net(fx, precision):
    from enerf_b200 import config as bcfg
    from enerf_b200.network_composite import Network
    bcfg.set_cfg(fx["cfg"])
    net = Network()
    net.load_state_dict(fx["state_dict"], strict=True)
    net = net.cuda().eval()
    net.precision = precision
    return net


@pytest.mark.parametrize("name", COMPOSITE_CASES)
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_composite_e2e_vs_reference_golden(name, precision):
    _cuda()
    fx = load_golden(name)
    net = _make_net(fx, precision)
    with torch.no_grad():
        out = net(_to_dev(fx["batch"]))
    torch.cuda.synchronize()
    tol, ztol = (5e-4, 2e-3) if precision == "fp32" else (2e-3, 4e-3)
    check_composite_outputs(out, fx["out"], tol=tol, ztol=ztol)
    from enerf_b200 import synthetic
    for k, ref in fx["out"].items():
        if k.startswith("rgb_"):     # north-star criterion carried over: |dPSNR| < 0.01 dB on a noisy target
            g = torch.Generator().manual_seed(3)
            target = (ref + 0.05 * torch.randn(ref.shape, generator=g)).clamp(0, 1)
            assert abs(synthetic.psnr(out[k].cpu(), target) - synthetic.psnr(ref, target)) < 0.01


def test_composite_overlap_and_device_rays_match():
    """Side-stream schedule off/on and batch rays vs on-device rays give the same frame (the kernels
    and their inputs are identical; only the ray generator's last-ulp differences remain)."""
    _cuda()
    fx = load_golden("c5_composite_2fg")
    net = _make_net(fx, "fp32")
    batch = _to_dev(fx["batch"])
    with torch.no_grad():
        net.overlap = True
        a = net(batch)
        net.overlap = False
        b = net(batch)
        c = net({k: v for k, v in batch.items() if not k.startswith("rays_")})
    torch.cuda.synchronize()
    for k in a:
        if a[k] is None:
            continue
        assert torch.equal(a[k], b[k]), k
        if a[k].dtype == torch.float32:
            assert (a[k] - c[k]).abs().max().item() <= 2e-5 * max(1.0, a[k].abs().max().item()), k


def test_windowed_volume_stages_match_full_grid():
    """cost_volume_window == the window of the full-grid cost volume (bit-exact: same arithmetic per
    voxel); depth_regress_window == depth_regress of the zero-padded probability volume."""
    _cuda()
    from enerf_b200 import capi, config as bcfg, synthetic
    cfg = bcfg.make_cfg(volume_planes=[8, 8])
    batch = synthetic.make_batch(64, 96, 3, cfg)
    levels = bcfg.snapshot(cfg)
    dev = "cuda"
    cam = torch.empty(capi.CAM_FLOATS, device=dev)
    scales = [(lv.im_feat_scale, lv.volume_scale, lv.render_scale) for lv in levels]
    capi.camera_setup(batch["src_exts"][0].cuda(), batch["src_ixts"][0].cuda(), batch["tar_ext"][0].cuda(), batch["tar_ixt"][0].cuda(),
                      batch["near_far"][0].cuda(), scales, cam)
    g = torch.Generator().manual_seed(5)
    for level, (h, w, C, hs, ws) in enumerate([(8, 12, 32, 16, 24), (32, 48, 16, 32, 48)]):
        D = 8
        feat = torch.randn(3, hs, ws, C, generator=g).cuda()
        ends, nf = torch.empty(2, h, w, device=dev), torch.empty(2, h, w, device=dev)
        lnf = torch.tensor([2.5, 4.5], device=dev)
        if level == 0:
            capi.depth_hypotheses_layer(lnf, None, None, None, h, w, D, True, ends, nf)
            assert torch.allclose(ends[0], torch.full((h, w), 2.5, device=dev)) and torch.allclose(ends[1], torch.full((h, w), 4.5, device=dev))
        else:
            pd, ps = 0.3 + 0.05 * torch.rand(8, 12, generator=g).cuda(), 0.02 * torch.rand(8, 12, generator=g).cuda()
            pnf = torch.stack([torch.full((8, 12), 0.4), torch.full((8, 12), 0.22)]).cuda()
            capi.depth_hypotheses_layer(None, pd, ps, pnf, h, w, D, False, ends, nf)
            e2, n2 = torch.empty_like(ends), torch.empty_like(nf)
            capi.depth_hypotheses(cam, pd, ps, pnf, h, w, D, False, e2, n2)
            assert torch.equal(ends, e2) and torch.equal(nf, n2)
        full = torch.empty(D, h, w, C, device=dev)
        capi.cost_volume(cam, level, feat, ends, D, h, w, level == 0, full)
        for win in ([2, 1, 8, 4], [0, 0, w, h], [w - 4, h - 4, 4, 4]):
            x, y, wc, hc = win
            part = torch.empty(D, hc, wc, C, device=dev)
            capi.cost_volume_window(cam, level, feat, ends, D, h, w, win, level == 0, part)
            assert torch.equal(part, full[:, y:y + hc, x:x + wc]), (level, win)
            prob = torch.randn(D, hc, wc, generator=g).cuda()
            padded = torch.zeros(D, h, w, device=dev)
            padded[:, y:y + hc, x:x + wc] = prob
            d0, s0, d1, s1 = (torch.empty(h, w, device=dev) for _ in range(4))
            capi.depth_regress(padded, ends, level == 0, d0, s0, None)
            capi.depth_regress_window(prob, win, ends, level == 0, d1, s1)
            assert torch.equal(d0, d1) and torch.equal(s0, s1), (level, win)
    with pytest.raises(ValueError, match="window"):
        capi.cost_volume_window(cam, 0, feat, ends, D, h, w, [w - 2, 0, 4, 4], False, full)


@pytest.mark.parametrize("L,ns", [(1, 2), (3, 2), (2, 1), (4, 8)])
def test_composite_layers_vs_oracle(L, ns):
    """enerf_composite_layers against raw2outputs_composite (oracle) on random layer samples, including
    overlapping, nested and empty windows."""
    _cuda()
    from enerf_b200 import capi, config as bcfg
    from oracle import enerf_oracle_composite as OC
    Hr, Wr = 40, 56
    g = torch.Generator().manual_seed(11 + L)
    boxes = [[4, 6, 32, 24], [20, 10, 30, 28], [0, 0, Wr, Hr], [50, 30, 0, 0]][:L]
    n_fg, n_tot = L * ns, L * ns + ns
    cfg = bcfg.composite_cfg(num_fg_layers=L, num=1, render_scale=[1.0], num_samples=[ns])
    batch = {"src_inps": torch.zeros(1, 2, 3, Hr, Wr), "bbox": torch.tensor([boxes], dtype=torch.float32)}
    layers, raw, z = [], torch.full((Hr * Wr, n_tot, 4), float("nan")), torch.full((Hr * Wr, n_tot), float("nan"))
    for l in range(L + 1):
        x, y, w, h = boxes[l] if l < L else (0, 0, Wr, Hr)
        lay = {"net_output": torch.rand(1, w * h, ns, 4, generator=g) * 2.0, "z_vals": 2.0 + 3.0 * torch.rand(1, w * h, ns, generator=g)}
        layers.append(lay)
        rv, zv = raw.view(Hr, Wr, n_tot, 4), z.view(Hr, Wr, n_tot)      # outside-window slots stay NaN: must never be read
        rv[y:y + h, x:x + w, l * ns:(l + 1) * ns] = lay["net_output"].view(h, w, ns, 4)
        zv[y:y + h, x:x + w, l * ns:(l + 1) * ns] = lay["z_vals"].view(h, w, ns)
    ref = OC.raw2outputs_composite(layers, batch, cfg, 0, L)
    dev = "cuda"
    rgb, dmap, wts = torch.empty(Hr * Wr, 3, device=dev), torch.empty(Hr * Wr, device=dev), torch.empty(Hr * Wr, n_tot, device=dev)
    net_out, z_vals = torch.empty(Hr * Wr, n_tot, 4, device=dev), torch.empty(Hr * Wr, n_fg, device=dev)
    idx = torch.empty(Hr * Wr, n_fg, device=dev, dtype=torch.int64) if L > 1 else None
    capi.composite_layers(raw.cuda(), z.cuda(), Hr, Wr, L, ns, ns, boxes, rgb, dmap, wts, net_out, idx, z_vals)
    out = {"rgb_level0": rgb[None], "depth_level0": dmap[None], "weights_level0": wts[None], "net_output_level0": net_out[None],
           "idx_level0": None if idx is None else idx[None], "z_vals_level0": z_vals[None]}
    check_composite_outputs(out, {f"{k}_level0": v for k, v in ref.items()}, tol=2e-6, ztol=2e-6)
    assert torch.equal(z_vals.cpu(), ref["z_vals"][0]) and torch.equal(net_out.cpu(), ref["net_output"][0])
