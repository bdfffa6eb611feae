"""CPU-side tests: parameter tree / state_dict compatibility, BN folding + packing, C-ABI symbols."""
import ctypes
import os
import re

import pytest
import torch

from enerf_b200 import capi, packing, synthetic
from enerf_b200 import config as bcfg
from _helpers import ROOT, load_golden


def _net(cfg):
    from enerf_b200.network import Network
    bcfg.set_cfg(cfg)
    torch.manual_seed(0)
    return Network()


def test_state_dict_keys_and_init_match_reference_golden():
    """tests/golden holds the reference's own state_dict (seed 0 + randomised BN): same keys, same
    shapes, and -- since the holders replay the reference's construction order -- same values."""
    for name in ("c1_nocascade", "c2_small_cascade"):
        fx = load_golden(name)
        net = _net(fx["cfg"])
        synthetic.randomize_bn_(net, seed=1)
        sd = net.state_dict()
        assert list(sd.keys()) == list(fx["state_dict"].keys())
        for k, v in fx["state_dict"].items():
            assert sd[k].shape == v.shape, k
            assert torch.equal(sd[k], v), k
        net.load_state_dict(fx["state_dict"], strict=True)
    assert len(load_golden("c2_small_cascade")["state_dict"]) == 184


def test_bn_folding_matches_conv_bn():
    fx = load_golden("c2_small_cascade")
    sd = fx["state_dict"]
    import torch.nn.functional as F
    x = torch.randn(1, 8, 9, 10, generator=torch.Generator().manual_seed(0))
    q = "feature_net.conv1.0"
    ref = F.relu(F.batch_norm(F.conv2d(x, sd[q + ".conv.weight"], None, 2, 2), sd[q + ".bn.running_mean"], sd[q + ".bn.running_var"],
                              sd[q + ".bn.weight"], sd[q + ".bn.bias"], False, 0.0, 1e-5))
    pk = packing.pack_feature_net(sd, "cpu")
    w, b = pk[4], pk[5]  # conv1.0: [25][8][16]
    w_t = w.view(5, 5, 8, 16).permute(3, 2, 0, 1)
    out = F.relu(F.conv2d(x, w_t, b, 2, 2))
    assert (out - ref).abs().max() < 1e-5
    # transposed conv: scale runs along dim 1
    x3 = torch.randn(1, 32, 2, 3, 4, generator=torch.Generator().manual_seed(1))
    q = "cost_reg_1.conv9"
    ref = F.batch_norm(F.conv_transpose3d(x3, sd[q + ".0.weight"], None, 2, 1, 1), sd[q + ".1.running_mean"], sd[q + ".1.running_var"],
                       sd[q + ".1.weight"], sd[q + ".1.bias"], False, 0.0, 1e-5)
    pr = packing.pack_cost_reg(sd, "cost_reg_1", 16, True, "cpu", True)
    w, b = pr[16], pr[17]  # conv9 after conv0..6 (14) + conv7 (2)
    w_t = w.view(3, 3, 3, 32, 16).permute(3, 4, 0, 1, 2)
    out = F.conv_transpose3d(x3, w_t, b, 2, 1, 1)
    assert (out - ref).abs().max() < 1e-5
    assert pr[-1].shape == (27, 8, 9) and len(pr) == 21
    assert packing.pack_cost_reg(sd, "cost_reg_0", 32, False, "cpu", False)[-1].shape == (27, 8, 1)
    pn = packing.pack_nerf(sd, "nerf_1", 11, True, "cpu")
    assert [tuple(t.shape) for t in pn[::2]] == [(4, 11), (33, 32), (32, 1), (32, 16), (24, 64), (64, 1), (103, 64), (64, 1)]


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports exactly what include/enerf_b200.h declares (no compute)."""
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    hdr = open(os.path.join(ROOT, "include", "enerf_b200.h")).read()
    declared = set(re.findall(r"ENERF_API\s+[\w\s\*]+?\b(enerf_\w+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    handle = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert capi.lib().enerf_abi_version() == 2
    # the EnerfCam struct size the Python side assumes
    m = re.search(r"typedef struct EnerfCam \{(.*?)\} EnerfCam;", hdr, re.S)
    dims = re.findall(r"float \w+((?:\[\w+\])+);", m.group(1))
    consts = {"ENERF_MAX_LEVELS": 4, "ENERF_MAX_VIEWS": 8}
    total = 0
    for d in dims:
        n = 1
        for tok in re.findall(r"\[(\w+)\]", d):
            n *= consts.get(tok, None) or int(tok)
        total += n
    assert total == capi.CAM_FLOATS


def test_no_cpu_fallback():
    cfg = bcfg.make_cfg(volume_planes=[8, 8])
    net = _net(cfg).eval()
    batch = synthetic.make_batch(64, 96, 3, cfg)
    with pytest.raises(ValueError, match="CUDA"):
        net(batch)
    with pytest.raises(ValueError):
        capi.ptr(torch.zeros(4))


def test_cfg_snapshot_and_ray_layout():
    cfg = bcfg.make_cfg(volume_planes=[48, 8], render_if=[False, True])
    lv = bcfg.snapshot(cfg)
    assert [l.planes for l in lv] == [48, 8] and [l.render_if for l in lv] == [False, True]
    assert lv[1].prev_depth_inv and not lv[1].depth_inv and lv[1].feat_ch == 8
    b = synthetic.make_batch(64, 96, 3, cfg)
    assert b["rays_1"].shape == (1, 64 * 96, 8) and b["rays_0"].shape == (1, 16 * 24, 8)
    r = b["rays_1"][0].view(64, 96, 8)
    assert r[5, 7, 6] == 7 and r[5, 7, 7] == 5  # u, v are pixel coordinates, row-major


def test_composite_state_dict_and_packing_match_reference_golden():
    """network_composite: the reference's own state_dict (seed 0 + randomised BN) is reproduced key by
    key and value by value; nerf_.NeRF weights embed into the 16-tensor ray-kernel layout with zero
    rows where nerf.NeRF reads the voxel feature."""
    from enerf_b200.network_composite import Network as CompositeNetwork
    for name in ("c5_composite_1fg", "c5_composite_2fg"):
        fx = load_golden(name)
        bcfg.set_cfg(fx["cfg"])
        torch.manual_seed(0)
        net = CompositeNetwork()
        synthetic.randomize_bn_(net, seed=1)
        sd = net.state_dict()
        assert list(sd.keys()) == list(fx["state_dict"].keys())
        for k, v in fx["state_dict"].items():
            assert sd[k].shape == v.shape and torch.equal(sd[k], v), k
        net.load_state_dict(fx["state_dict"], strict=True)
        with pytest.raises(ValueError, match="CUDA"):
            net.eval()(fx["batch"])
    sd = load_golden("c5_composite_1fg")["state_dict"]
    pn = packing.pack_nerf_novox(sd, "nerf_1_bg", 11, False, "cpu")
    assert pn[0] is None and pn[1] is None
    assert tuple(pn[8].shape) == (24, 64) and tuple(pn[12].shape) == (103, 64)
    assert pn[8][:8].abs().max() == 0 and pn[12][64:72].abs().max() == 0
    assert torch.equal(pn[8][8:], sd["nerf_1_bg.lr0.0.weight"].t()) and torch.equal(pn[12][72:], sd["nerf_1_bg.color.0.weight"].t()[64:])


def test_folded_tc_conv_packing_is_a_shifted_sum_of_partial_convolutions():
    """packing.pack_tc_conv(fold_kx=True) (the layout csrc/tc_conv.cu consumes for the 3-D layers with 8
    or 1 output channels): un-packing it and evaluating  out[x] = sum_kx P[x + kx][kx]  with
    P[x][kx] = sum_{kz,ky,c} in[z+kz-1, y+ky-1, x-1][c] * W[kz][ky][kx][c]  reproduces conv3d -- the
    identity the kernel's epilogue relies on (two lane shifts)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    cin, cout, D, H, W = 16, 8, 3, 5, 7
    w = torch.randn(cout, cin, 3, 3, 3, generator=g)
    x = torch.randn(1, cin, D, H, W, generator=g)
    assert packing.tc_fold_kx(3, 3, 1, cout) and not packing.tc_fold_kx(3, 3, 2, cout) and not packing.tc_fold_kx(3, 3, 1, 16)
    assert packing.tc_fold_kx(3, 3, 1, 1, single=True) and not packing.tc_fold_kx(1, 3, 1, 16) and not packing.tc_fold_kx(1, 5, 2, 8)
    rule = packing.FOLD_RULE
    try:                                    # the rule levels (mirrors csrc/tc_conv.cu::tc_fold_rule)
        packing.FOLD_RULE = 0
        assert not packing.tc_fold_kx(3, 3, 1, 9, head=True) and not packing.tc_fold_kx(1, 3, 1, 8)
        packing.FOLD_RULE = 1
        assert packing.tc_fold_kx(3, 3, 1, 9, head=True) and not packing.tc_fold_kx(1, 3, 1, 8)
        packing.FOLD_RULE = 2
        assert packing.tc_fold_kx(3, 3, 1, 9, head=True) and packing.tc_fold_kx(1, 3, 1, 8, cin=32) and not packing.tc_fold_kx(1, 3, 1, 8, cin=8)
    finally:
        packing.FOLD_RULE = rule
    taps = packing._taps_cin_cout(w)                                   # [27][cin][cout], taps ordered (kz,ky,kx)
    packed = packing.pack_tc_conv(taps, fold_kx=True)
    N = 32                                                             # 3 * 8 = 24 padded to 16-multiples
    assert packed.numel() == (cin // 8) * 9 * 2 * N * 4
    # [cin/8][tap_zy][2][N][4] -> [tap_zy][cin][N]
    wz = packed.view(cin // 8, 9, 2, N, 4).permute(1, 0, 2, 4, 3).reshape(9, cin, N)
    assert wz[:, :, 24:].abs().max() == 0
    ref_w = packing.tf32_round(taps.float()).view(9, 3, cin, cout)
    assert torch.equal(wz[:, :, :24].reshape(9, cin, 3, cout).permute(0, 2, 1, 3), ref_w)
    # partial convolutions over (kz,ky) only, evaluated at input column x-1 (zero padded), then the shifted sum
    xp = F.pad(x[0].permute(1, 2, 3, 0), (0, 0, 1, 3, 1, 1, 1, 1))     # (D+2, H+2, W+4, cin): x index i <-> column i-1
    P = torch.zeros(D, H, W + 2, 3, cout)
    for kz in range(3):
        for ky in range(3):
            blk = xp[kz:kz + D, ky:ky + H, 0:W + 2]                    # column j holds in[.., .., j-1]
            P += torch.einsum("zyxc,ckn->zyxkn", blk, wz[kz * 3 + ky, :, :24].reshape(cin, 3, cout))
    out = sum(P[:, :, kx:kx + W, kx] for kx in range(3))               # out[x] = sum_kx P[x+kx][kx]
    ref = F.conv3d(x, packing.tf32_round(w), None, 1, 1)[0].permute(1, 2, 3, 0)
    assert (out - ref).abs().max().item() < 1e-4


def test_band_rows_cover_the_receptive_field():
    """Network._band_rows: the volume rows a rank regularises = its rows + the regulariser's halo, clamped to the grid and
    aligned to the U-Net's stride; misaligned bands are refused (the bit-identity argument needs aligned crops)."""
    import pytest
    from enerf_b200 import config as bcfg
    from enerf_b200.config import snapshot
    from enerf_b200.network import Network
    cfg = bcfg.set_cfg(bcfg.make_cfg(volume_planes=[48, 8], render_if=[False, True]))
    net = Network()
    lv = snapshot(cfg)[1]
    H, h = 512, 256
    for world in (2, 4, 8):
        rows = H // world
        for rank in range(world):
            net.ray_rows = (rank * rows, (rank + 1) * rows)
            y0, y1 = net._band_rows(lv, H, h, True)
            v0, v1 = rank * rows // 2, (rank + 1) * rows // 2
            assert y0 == max(0, v0 - 32) and y1 == min(h, v1 + 32) and y0 % 8 == 0 and y1 % 8 == 0
            assert (v0 - y0 >= 31 or y0 == 0) and (y1 - v1 >= 31 or y1 == h)      # RF 30 rows + 1 for the bilinear / trilinear taps
    net.ray_rows = (0, 40)      # 20 volume rows: not a multiple of the 3-level U-Net's stride 8
    with pytest.raises(ValueError):
        net._band_rows(lv, H, h, True)
    net.ray_rows = None


def test_fastdiv_constants_divide_exactly():
    """Divisions by run-time extents inside the kernels (cost volume: voxel -> (d, y, x)) use multiply-high + shift constants made on
    the host (csrc/common.cuh FastDiv): exact for every divisor >= 1 and every n < 2^31."""
    import random
    from enerf_b200 import capi
    f = capi.lib().enerf_fastdiv_check
    rnd = random.Random(7)
    for d in list(range(1, 70)) + [80, 160, 320, 640, 1088, 1920, 4095, 4096, 4097, 65537, (1 << 20) + 3, (1 << 30) + 1, (1 << 31) - 1]:
        for n in [0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 31) - 1, (1 << 31) - 2] + [rnd.randrange(0, 1 << 31) for _ in range(200)]:
            if 0 <= n < (1 << 31):
                assert f(d, n) == n // d, (d, n)


def test_tf32_rounding_by_integer_add_is_round_to_nearest_ties_away():
    """csrc/tc.cuh::to_tf32 and packing.tf32_round round an fp32 value to TF32 by adding half a TF32 ulp (0x1000) to the bit pattern;
    the tensor core then ignores the low 13 mantissa bits.  Checked here in VALUE space against an independent float64
    construction of round-to-nearest, ties away from zero, on a 10-bit mantissa (= cvt.rna.tf32.f32) -- including ties, values
    that carry into the next exponent, and subnormal-free small magnitudes."""
    import numpy as np
    from enerf_b200 import packing
    rnd = np.random.default_rng(11)
    x = np.concatenate([rnd.standard_normal(20000).astype(np.float32) * np.float32(10.0) ** rnd.integers(-20, 20, 20000).astype(np.float32),
                        np.float32([1.0, -1.0, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, -(1.0 + 2.0 ** -11), 2.0 - 2.0 ** -12, 0.0,
                                    1.9999999, 3.0e38 * 0.5, 1.2e-38 * 4])])
    got = packing.tf32_round(torch.from_numpy(x)).numpy().astype(np.float64)
    xd = np.abs(x.astype(np.float64))
    nz = xd > 0
    e = np.floor(np.log2(xd, where=nz, out=np.zeros_like(xd)))
    ulp = np.exp2(e - 10)
    lo = np.floor(xd / ulp) * ulp
    hi = lo + ulp
    ref = np.where(xd - lo < hi - xd, lo, hi)          # ties (equal distances) go to hi = away from zero
    ref = np.where(nz, np.sign(x.astype(np.float64)) * ref, 0.0)
    assert np.array_equal(got, ref)


def test_committed_bench_line_keeps_the_contract():
    """The last measured bench.py line of the round (profiles/r2_bench_1gpu_final.json, written on the B200 box) carries every key the
    measurement contract names, with consistent values: value = frames per step / time per step, roofline.frac = achieved / peak,
    e2e measured with real host <-> device copies, clocks without a thermal / hardware slow-down."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_bench_1gpu_final.json")
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    frames = d["config"]["frames_in_flight_per_gpu"]
    assert abs(d["value"] - frames / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"] * 1.05
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"]
