"""GPU parity tests (run on the B200 with `-m gpu`): every C-ABI stage against the CPU oracle on
the oracle's own inputs, and the drop-in Network end to end against (a) the oracle and (b) the
reference-minted golden outputs.  Tolerances are absolute fp32 bounds, stated per quantity:

  stage outputs (fed with oracle inputs)    : 1e-4 * max(1, |ref|_max)   (fp32 conv/gather round-off)
  end-to-end rgb                            : 5e-4 abs and |dPSNR| < 0.01 dB   (north-star criterion)
  end-to-end depth / depth_mvs / std / wts  : 2e-3 * max(1, |ref|_max)   (cascade amplifies round-off
                                              of the level-0 softmax into level-1 hypotheses)
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

STAGE_TOL = 1e-4
# tensor-core (TF32 operand) stages: the reference's own sensitivity to TF32 MLP operands is
# max|d rgb| 2.8e-4; bound = 2e-3 abs on rgb / weights, 2e-3 rel on depth
TC_STAGE_TOL = 2e-3
E2E_TOL = {"rgb": 5e-4, "depth": 2e-3, "weights": 2e-3, "depth_mvs": 2e-3, "std": 2e-3}
E2E_TOL_TF32 = {"rgb": 2e-3, "depth": 4e-3, "weights": 4e-3, "depth_mvs": 2e-3, "std": 2e-3}


def _check_report(rep, e2e_tol=E2E_TOL):
    bad = []
    for k, v in rep.items():
        if "max_abs" not in v:
            if abs(v["delta_psnr"]) >= 0.01:
                bad.append((k, v))
            continue
        if v["nan"]:
            bad.append((k, v))
            continue
        if k.startswith("e2e."):
            tol = e2e_tol[k[4:].split("_level")[0]] * max(1.0, v["ref_max"])
        elif k.split(".")[0] in ("render_rays_tc", "feature_net_tc", "cost_reg_tc"):
            tol = TC_STAGE_TOL * max(1.0, v["ref_max"])
        else:
            tol = STAGE_TOL * max(1.0, v["ref_max"])
        if v["max_abs"] > tol:
            bad.append((k, v, tol))
    assert not bad, "parity failures:\n" + "\n".join(map(str, bad))


@pytest.fixture(scope="module")
def harness():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import stage_harness
    return stage_harness


def test_stages_and_e2e_vs_oracle_on_golden_inputs(harness, golden):
    rep, out = harness.stage_report(golden["state_dict"], golden["cfg"], golden["batch"], human=golden.get("human", False))
    _check_report(rep)
    # (b) directly against what the unmodified reference produced
    for k, ref in golden["out"].items():
        tol = E2E_TOL[k.split("_level")[0]] * max(1.0, ref.abs().max().item())
        e = (out[k].cpu() - ref).abs().max().item()
        assert e <= tol, f"{k} vs reference golden: {e} > {tol}"


@pytest.mark.parametrize("H,W,S,planes,render_if", [(128, 160, 3, (16, 8), (False, True)), (96, 128, 2, (8, 8), (True, True)),
                                                    (64, 64, 5, (8, 8), (False, True))])
def test_stages_and_e2e_vs_oracle_synthetic(harness, H, W, S, planes, render_if):
    from enerf_b200 import config as bcfg
    cfg = bcfg.make_cfg(volume_planes=list(planes), render_if=list(render_if))
    sd, batch = harness.make_case(H, W, S, cfg, seed=11)
    rep, _ = harness.stage_report(sd, cfg, batch)
    _check_report(rep)


def test_white_bkgd_and_no_viewdir(harness):
    from enerf_b200 import config as bcfg
    cfg = bcfg.make_cfg(volume_planes=[8, 8], render_if=[False, True], white_bkgd=True, viewdir_agg=False)
    sd, batch = harness.make_case(64, 96, 3, cfg, seed=7)
    rep, _ = harness.stage_report(sd, cfg, batch)
    _check_report(rep)


def test_ray_subset_matches_full_frame(harness):
    """Rays are independent: rendering a row band gives exactly the rows of the full frame
    (the property the multi-GPU sharding relies on)."""
    from enerf_b200 import config as bcfg
    from enerf_b200.network import Network
    cfg = bcfg.make_cfg(volume_planes=[8, 8], render_if=[False, True])
    sd, batch = harness.make_case(64, 96, 3, cfg, seed=3)
    net = Network()
    net.load_state_dict(sd)
    net = net.cuda().eval()
    gb = {k: v.cuda() for k, v in batch.items()}
    with torch.no_grad():
        full = net(gb)
        band = dict(gb)
        band["rays_1"] = gb["rays_1"][:, 96 * 16: 96 * 48].contiguous()
        part = net(band)
    for k in ("rgb_level1", "depth_level1", "weights_level1"):
        assert torch.equal(part[k], full[k][:, 96 * 16: 96 * 48]), k


def test_errors_are_loud(harness):
    from enerf_b200 import config as bcfg
    from enerf_b200.network import Network
    cfg = bcfg.make_cfg(volume_planes=[8, 8], render_if=[False, True])
    sd, batch = harness.make_case(64, 96, 3, cfg, seed=3)
    net = Network().cuda().eval()
    with pytest.raises(ValueError):
        net(batch)  # CPU tensors: no fallback
    bad = {k: v.cuda() for k, v in batch.items()}
    bad["src_inps"] = bad["src_inps"][..., :60, :]  # volume not divisible
    with pytest.raises(ValueError):
        net(bad)
    with pytest.raises(NotImplementedError):
        net.train()(bad)


def _tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) keeping 10 mantissa bits."""
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("K,N", [(8, 16), (16, 32), (24, 64), (32, 16), (88, 64), (104, 64), (128, 256)])
def test_tcgen05_selftest_gemm(K, N):
    """Pins the UMMA shared-memory / instruction descriptor encodings and the TMEM row mapping
    of csrc/tc.cuh: D = A B^T with TF32 operands must match an fp64 product of the rounded inputs."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enerf_b200 import capi
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(N, K, generator=g)
    D = torch.full((128, N), float("nan"), device="cuda")
    capi.tc_selftest(A.cuda(), B.cuda(), D)
    torch.cuda.synchronize()
    ref = (_tf32_rna(A).double() @ _tf32_rna(B).double().t()).float()
    err = (D.cpu() - ref).abs().max().item()
    assert err < 1e-4 * max(1.0, ref.abs().max().item()), f"K={K} N={N}: max abs err {err}"


@pytest.mark.parametrize("Kf", [8, 16, 32])
def test_tcgen05_swizzled_tma_operand_with_row_offsets(Kf):
    """The convention the TMA-fed convolution kernel rests on (csrc/tma.cuh, tc_conv2.cu): an operand tile written by ONE
    TMA box with SWIZZLE_{32,64,128}B is read by tcgen05.mma through a K-major swizzled descriptor whose start address is
    advanced by an arbitrary number of whole rows (a filter tap) and by 32 bytes per K-step, base_offset = 0 -- the swizzle
    XOR is a function of the absolute shared-memory address on both sides.  One-hot B: D must EQUAL the shifted rows of A."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enerf_b200 import capi
    g = torch.Generator().manual_seed(Kf)
    A = _tf32_rna(torch.randn(256, Kf, generator=g)).cuda()
    N = max(16, Kf)
    B = torch.zeros(N, Kf)
    B[torch.arange(Kf), torch.arange(Kf)] = 1.0
    B = B.cuda()
    for row_off in (0, 1, 3, 8, 13, 34, 70, 105):
        D = torch.full((128, N), float("nan"), device="cuda")
        capi.tc_swz_selftest(A, B, D, row_off, 0)
        torch.cuda.synchronize()
        assert torch.equal(D[:, :Kf], A[row_off:row_off + 128]), f"Kf={Kf} row_off={row_off}"
        assert not D[:, Kf:].any()


def test_e2e_tf32_tensor_core_path(harness, golden):
    """Default precision: the ray-stage MLP runs on tcgen05 (TF32 operands).  Same oracle, wider
    element-wise bound, same |dPSNR| < 0.01 dB criterion."""
    rep, out = harness.stage_report(golden["state_dict"], golden["cfg"], golden["batch"], precision="tf32", human=golden.get("human", False))
    _check_report({k: v for k, v in rep.items() if k.startswith("e2e.")}, E2E_TOL_TF32)


TC_CONV_CASES = [
    # kind, KD, KH, cin, cout, mode, relu, (D, H, W)
    (0, 3, 3, 16, 8, 0, 1, (8, 32, 48)),     # CostRegNet conv0
    (0, 3, 3, 32, 8, 0, 1, (4, 16, 40)),     # MinCostRegNet conv0
    (0, 3, 3, 16, 16, 0, 1, (4, 16, 24)),    # conv2
    (0, 3, 3, 32, 32, 0, 1, (2, 8, 12)),     # conv4
    (0, 3, 3, 64, 64, 0, 1, (1, 4, 6)),      # conv6 (8 K-stages through the 2-deep ring)
    (0, 3, 3, 8, 9, 1, 0, (8, 32, 48)),      # head: feat_conv + depth_conv
    (0, 3, 3, 8, 1, 3, 0, (8, 16, 24)),      # depth_conv only
    (0, 1, 3, 32, 8, 0, 0, (3, 64, 96)),     # FeatureNet smooth0 (2-D: images on the depth axis)
    (0, 1, 3, 32, 32, 0, 1, (2, 16, 24)),    # conv2.1
    (0, 1, 1, 32, 32, 0, 0, (3, 16, 24)),    # toplayer 1x1
    (1, 3, 3, 16, 8, 2, 0, (4, 16, 24)),     # conv11 (transposed)
    (1, 3, 3, 32, 16, 2, 0, (2, 8, 12)),     # conv9
    (1, 3, 3, 64, 32, 2, 0, (1, 4, 6)),      # conv7
    (0, 3, 3, 8, 16, 0, 1, (8, 32, 48), 2),  # conv1 (stride 2: phase-tile staging)
    (0, 3, 3, 16, 32, 0, 1, (4, 16, 24), 2),  # conv3
    (0, 3, 3, 32, 64, 0, 1, (2, 8, 12), 2),  # conv5
    (0, 1, 5, 8, 16, 0, 1, (3, 64, 96), 2),  # FeatureNet conv1.0 (5x5 stride 2)
    (0, 1, 5, 16, 32, 0, 1, (2, 32, 48), 2),  # conv2.0
    (0, 3, 3, 8, 8, 0, 1, (5, 20, 44)),      # ragged extents: partial tiles in every dimension (folded kx, row exchange)
    (0, 1, 3, 8, 8, 0, 1, (3, 50, 70)),      # conv0.1-like, ragged
    (0, 3, 3, 8, 9, 1, 0, (4, 12, 20)),      # head, ragged
    # full-size layers: many tiles per persistent CTA (ring wrap-around, both accumulator stages, K-block pipelining)
    (0, 1, 3, 32, 8, 0, 0, (3, 512, 640)),   # smooth0 at the headline size
    (0, 3, 3, 16, 8, 0, 1, (8, 256, 320)),   # CostRegNet conv0 (folded) at the headline size
    (0, 3, 3, 32, 8, 0, 1, (16, 64, 80)),    # MinCostRegNet conv0: 32 channels = two 16-channel K-blocks
    (0, 3, 3, 8, 9, 1, 0, (8, 128, 160)),    # head
    (1, 3, 3, 16, 8, 2, 0, (4, 64, 80)),     # conv11 (transposed), many tiles
    (0, 3, 3, 8, 16, 0, 1, (8, 128, 160), 2),   # conv1 (stride 2) with many tiles
    (0, 1, 5, 8, 16, 0, 1, (3, 256, 320), 2),   # FeatureNet conv1.0 (5x5 stride 2) with many tiles
]

# which kernel runs the layer: csrc/tc_conv.cu ("v1"), or the persistent TMA-fed csrc/tc_conv2.cu where it is eligible
# (auto = the shipped policy: 2 persistent CTAs per SM x 2 MMA-issuing warps; the 1-CTA variants)
TC_CONV_IMPLS = {"v1": dict(impl=1), "auto": dict(impl=0), "1cta1mma": dict(impl=0, nmma=1, ctas_per_sm=1), "1cta2mma": dict(impl=0, nmma=2, ctas_per_sm=1),
                 "nos2": dict(impl=3)}    # impl 3: the stride-2 layers stay on csrc/tc_conv.cu (default: TMA boxes with element stride 2)


@pytest.mark.parametrize("impl", list(TC_CONV_IMPLS))
@pytest.mark.parametrize("case", TC_CONV_CASES)
def test_tc_conv_layer(case, impl):
    _run_tc_conv_layer(case, impl)


@pytest.mark.parametrize("rule", [0, 1])
@pytest.mark.parametrize("impl", ["v1", "auto"])
@pytest.mark.parametrize("case", [c for c in TC_CONV_CASES if (c[1] == 1 and c[2] == 3 and c[4] == 8) or c[5] == 1])
def test_tc_conv_layer_fold_rule_levels(case, impl, rule):
    """The layers the kx-fold rule levels 1 / 2 add (feat + prob head; 3x3 2-D layers with 8 output channels) under the
    narrower rules too (enerf_tc_conv_fold_rule; the shipped level 2 is what test_tc_conv_layer runs)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enerf_b200 import capi
    capi.tc_conv_fold_rule(rule)
    try:
        _run_tc_conv_layer(case, impl)
    finally:
        capi.tc_conv_fold_rule(2)


def _run_tc_conv_layer(case, impl):
    kind, KD, KH, cin, cout, mode, relu, dims = case[:8]
    stride = case[8] if len(case) > 8 else 1
    if impl == "nos2" and stride == 1:
        pytest.skip("impl 3 differs from auto on the stride-2 layers only")
    if impl in ("1cta1mma", "1cta2mma") and dims[1] * dims[2] < 64 * 80:
        pytest.skip("variant exercised on the larger cases")
    """tcgen05 implicit-GEMM convolution vs torch's fp32 conv on the same (TF32-rounded) operands.
    Bound: 2e-3 * max|ref| (TF32 operand rounding of activations; weights are pre-rounded)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import torch.nn.functional as F
    from enerf_b200 import capi, packing
    D, H, W = dims
    g = torch.Generator().manual_seed(cin * 100 + cout + kind)
    x = torch.randn(1, cin, D, H, W, generator=g)
    if kind == 0:
        w = torch.randn(cout, cin, KD, KH, KH, generator=g) / (cin * KD * KH * KH) ** 0.5
        b = torch.randn(cout, generator=g) * 0.1 if mode == 0 else None
        ref = F.conv3d(x, w, b, (stride if KD > 1 else 1, stride, stride), (KD // 2, KH // 2, KH // 2))
        if relu:
            ref = F.relu(ref)
        wp = packing.pack_tc_conv(packing._taps_cin_cout(w), fold_kx=packing.tc_fold_kx(KD, KH, stride, cout, single=(mode == 3), head=(mode == 1), cin=cin)).cuda()
        skip = None
    else:
        w = torch.randn(cin, cout, 3, 3, 3, generator=g) / (cin * 27 / 8) ** 0.5
        b = torch.randn(cout, generator=g) * 0.1
        skip_t = torch.randn(1, cout, 2 * D, 2 * H, 2 * W, generator=g)
        ref = skip_t + F.conv_transpose3d(x, w, b, stride=2, padding=1, output_padding=1)
        wp = packing.pack_tc_deconv(w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout)).cuda()
        skip = skip_t[0].permute(1, 2, 3, 0).contiguous().cuda()
    xin = x[0].permute(1, 2, 3, 0).contiguous().cuda()
    Do, Ho, Wo = ref.shape[2:]
    out2 = None
    if mode == 1:
        out = torch.full((Do, Ho, Wo, 8), float("nan"), device="cuda")
        out2 = torch.full((Do, Ho, Wo), float("nan"), device="cuda")
    elif mode == 3:
        out = torch.full((Do, Ho, Wo), float("nan"), device="cuda")
    else:
        out = torch.full((Do, Ho, Wo, cout), float("nan"), device="cuda")
    capi.tc_conv2_tune(**TC_CONV_IMPLS[impl])
    try:
        capi.tc_conv(kind, KD, KH, cout, mode, relu, xin, wp, b.cuda() if b is not None else None, skip, out, out2,
                     out_cstride=(8 if mode == 1 else cout), stride=stride)
        torch.cuda.synchronize()
    finally:
        capi.tc_conv2_tune()
    refc = ref[0].permute(1, 2, 3, 0)
    if mode == 1:
        got = torch.cat([out.cpu(), out2.cpu()[..., None]], dim=-1)
    elif mode == 3:
        got = out.cpu()[..., None]
    else:
        got = out.cpu()
    err = (got - refc).abs().max().item()
    assert not torch.isnan(got).any()
    assert err < 2e-3 * max(1.0, refc.abs().max().item()), f"max abs err {err} (ref max {refc.abs().max().item()})"


@pytest.mark.parametrize("H,W,S", [(64, 96, 2), (512, 640, 3)])
def test_feature_net_fused_lateral_is_bit_identical(H, W, S):
    """FeatureNet on the tensor-core path: lat0 computed inside smooth0's producer warps (tc_conv2.cu, PROD = 1) against the
    separate lateral kernel + smooth0, and both tcgen05 convolution kernels against each other: identical arithmetic in
    identical order, so level_2 / level_1 / level_0 features must be EQUAL."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enerf_b200 import capi, packing
    from enerf_b200 import config as bcfg
    import stage_harness
    cfg = bcfg.make_cfg(volume_planes=[8, 8], render_if=[False, True])
    sd, _ = stage_harness.make_case(64, 96, 2, cfg, seed=5)
    g = torch.Generator().manual_seed(H + S)
    src = (2 * torch.rand(S, 3, H, W, generator=g) - 1).cuda()
    pk = packing.pack_feature_net(sd, torch.device("cuda"), tensor_cores=True)
    ws = torch.empty(capi.feature_net_workspace_bytes(S, H, W) // 4, device="cuda")
    outs = {}
    try:
        for name, impl, fuse in (("v1", 1, False), ("v2", 0, False), ("v2_fused", 0, True)):
            capi.tc_conv2_tune(impl=impl)
            capi.tc_conv2_fuse_lateral(fuse)
            f0 = torch.full((S, H // 4, W // 4, 32), float("nan"), device="cuda")
            f1 = torch.full((S, H // 2, W // 2, 16), float("nan"), device="cuda")
            f2 = torch.full((S, H, W, 8), float("nan"), device="cuda")
            capi.feature_net(pk, src, f0, f1, f2, ws, tensor_cores=True)
            torch.cuda.synchronize()
            outs[name] = (f0, f1, f2)
    finally:
        capi.tc_conv2_tune()
        capi.tc_conv2_fuse_lateral(True)
    for name in ("v2", "v2_fused"):
        for lvl in range(3):
            assert not torch.isnan(outs[name][lvl]).any(), (name, lvl)
            assert torch.equal(outs[name][lvl], outs["v1"][lvl]), (name, lvl, (outs[name][lvl] - outs["v1"][lvl]).abs().max().item())


@pytest.mark.parametrize("H,W,S", [(64, 96, 2), (512, 640, 3)])
def test_feature_net_packed_output_equals_the_pack_kernel(H, W, S):
    """enerf_feature_net_packed: the (feature | rgb | 0) records written by the fused lat0 + smooth0 launch's epilogue (tensor-core
    path), and by the internal pack kernel when there is no fused launch (fusion off, FP32 mode), must EQUAL what
    enerf_pack_img_feat builds from feat_l2 and the source images; feat_l2 itself must be unchanged by the extra output."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enerf_b200 import capi, packing
    from enerf_b200 import config as bcfg
    import stage_harness
    cfg = bcfg.make_cfg(volume_planes=[8, 8], render_if=[False, True])
    sd, _ = stage_harness.make_case(64, 96, 2, cfg, seed=5)
    g = torch.Generator().manual_seed(H + 3 * S)
    src = (2 * torch.rand(S, 3, H, W, generator=g) - 1).cuda()
    ws = torch.empty(capi.feature_net_workspace_bytes(S, H, W) // 4, device="cuda")
    try:
        for name, tcs, fuse in (("tf32_fused", True, True), ("tf32_separate", True, False), ("fp32", False, True)):
            pk = packing.pack_feature_net(sd, torch.device("cuda"), tensor_cores=tcs)
            capi.tc_conv2_fuse_lateral(fuse)
            f0 = torch.empty((S, H // 4, W // 4, 32), device="cuda")
            f1 = torch.empty((S, H // 2, W // 2, 16), device="cuda")
            f2_plain = torch.full((S, H, W, 8), float("nan"), device="cuda")
            capi.feature_net(pk, src, f0, f1, f2_plain, ws, tensor_cores=tcs)
            ref = torch.full((S, H, W, 12), float("nan"), device="cuda")
            capi.pack_img_feat(f2_plain, src, ref)
            f2 = torch.full((S, H, W, 8), float("nan"), device="cuda")
            img = torch.full((S, H, W, 12), float("nan"), device="cuda")
            capi.feature_net(pk, src, f0, f1, f2, ws, tensor_cores=tcs, img_feat_rgb=img)
            torch.cuda.synchronize()
            assert not torch.isnan(img).any(), name
            assert torch.equal(f2, f2_plain), name
            assert torch.equal(img, ref), (name, (img - ref).abs().max().item())
    finally:
        capi.tc_conv2_fuse_lateral(True)


def test_device_ray_generation_matches_data_layer(harness):
    """enerf_generate_rays vs the reference's numpy builder (restated in synthetic.full_frame_rays,
    itself cross-checked against lib/datasets/enerf_utils.py in oracle/make_golden.py)."""
    from enerf_b200 import capi, synthetic
    from enerf_b200 import config as bcfg
    cfg = bcfg.make_cfg(volume_planes=[8, 8], render_if=[True, True])
    batch = synthetic.make_batch(64, 96, 3, cfg, seed=5)
    te, ti = batch["tar_ext"][0].cuda().contiguous(), batch["tar_ixt"][0].cuda().contiguous()
    for scale, Hs, Ws in ((1.0, 64, 96), (0.25, 16, 24)):
        out = torch.empty((Hs * Ws, 8), device="cuda")
        capi.generate_rays(te, ti, scale, Ws, 0, Hs, out)
        ref = torch.from_numpy(synthetic.full_frame_rays(batch["tar_ext"][0].numpy(), batch["tar_ixt"][0].numpy(), 64, 96, scale))
        assert (out.cpu() - ref).abs().max().item() <= 1e-6 * max(1.0, ref.abs().max().item())
        band = torch.empty((4 * Ws, 8), device="cuda")
        capi.generate_rays(te, ti, scale, Ws, 3, 4, band)
        assert torch.equal(band, out[3 * Ws:7 * Ws])


def test_forward_without_rays_graph_and_streamed(harness):
    """(a) a batch without rays_{i} renders identically (rays generated on device); (b) CUDA-graph
    replay and (c) the 3-stream StreamedRenderer return exactly what Network.forward returns."""
    from enerf_b200 import config as bcfg
    from enerf_b200.network import Network
    from enerf_b200.pipeline import GraphedNetwork, StreamedRenderer
    cfg = bcfg.make_cfg(volume_planes=[8, 8], render_if=[False, True])
    sd, batch = harness.make_case(64, 96, 3, cfg, seed=3)
    net = Network()
    net.load_state_dict(sd)
    net = net.cuda().eval()
    gb = {k: v.cuda() for k, v in batch.items()}
    norays = {k: v for k, v in gb.items() if not k.startswith("rays_")}
    net.precision = "fp32"      # exact mode: a 1-ulp change of the ray directions stays a 1e-5 change of the output
    with torch.no_grad():
        full = {k: v.clone() for k, v in net(gb).items()}
        gen32 = net(norays)
    for k in full:
        assert (gen32[k] - full[k]).abs().max().item() <= 2e-5 * max(1.0, full[k].abs().max().item()), k
    net.precision = "tf32"
    with torch.no_grad():
        gen = {k: v.clone() for k, v in net(norays).items()}
    g = GraphedNetwork(net, norays)
    out = g(norays)
    torch.cuda.synchronize()
    for k in gen:
        assert torch.equal(out[k], gen[k]), k
    host = {k: v.cpu().pin_memory() for k, v in norays.items()}
    got = []
    sr = StreamedRenderer(net, host, torch.device("cuda"), depth=3)
    sr.render([host] * 5, lambda i, o: got.append({k: v.clone() for k, v in o.items()}))
    assert len(got) == 5
    for o in got:
        for k in gen:
            assert torch.equal(o[k], gen[k].cpu()), k


def test_streamed_distinct_frames_do_not_share_scratch(harness):
    """Frames in flight must not share the Network's scratch (camera struct, workspaces, volumes):
    DIFFERENT frames (other images, other cameras) streamed with depth 3 / replayed concurrently as
    graph replicas must each equal their own eager forward (the captures used to alias)."""
    from enerf_b200 import config as bcfg
    from enerf_b200.network import Network
    from enerf_b200.pipeline import GraphedNetwork, StreamedRenderer
    cfg = bcfg.make_cfg(volume_planes=[16, 8], render_if=[False, True])
    sd, _ = harness.make_case(128, 160, 3, cfg, seed=3)
    net = Network()
    net.load_state_dict(sd)
    net = net.cuda().eval()
    from enerf_b200 import synthetic
    hosts, want = [], []
    for i in range(6):
        b = synthetic.make_batch(128, 160, 3, cfg, seed=20 + 7 * i)
        b = {k: v for k, v in b.items() if not k.startswith("rays_")}
        hosts.append({k: v.pin_memory() for k, v in b.items()})
        with torch.no_grad():
            want.append({k: v.clone().cpu() for k, v in net({k: v.cuda() for k, v in b.items()}).items()})
    assert (want[0]["rgb_level1"] - want[1]["rgb_level1"]).abs().max().item() > 1e-2     # the frames really differ
    got = {}
    sr = StreamedRenderer(net, hosts[0], torch.device("cuda"), depth=3)
    for _ in range(3):                                   # several passes: any aliasing shows up as a race
        sr.render(hosts, lambda i, o: got.__setitem__(i, {k: v.clone() for k, v in o.items()}))
        for i in range(6):
            for k in want[i]:
                assert torch.equal(got[i][k], want[i][k]), (i, k)
    # graph replicas replayed concurrently on their own streams (what bench.py does)
    reps = []
    for j in range(3):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            reps.append((GraphedNetwork(net, {k: v.cuda() for k, v in hosts[j].items()}), st))
    torch.cuda.synchronize()
    ptrs = [r.scratch["cam"].data_ptr() for r, _ in reps]
    assert len(set(ptrs)) == len(ptrs), "graph replicas share the camera scratch"
    for _ in range(5):
        outs = []
        for j, (r, st) in enumerate(reps):
            with torch.cuda.stream(st):
                r.load({k: v.cuda() for k, v in hosts[j + 3].items()}, non_blocking=False)
                outs.append(r.replay())
        torch.cuda.synchronize()
        for j, o in enumerate(outs):
            for k in want[j + 3]:
                assert torch.equal(o[k].cpu(), want[j + 3][k]), (j, k)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_band_sharding_is_bit_identical(harness, world, precision):
    """The row-band multi-GPU layout (level-1 cost volume / CostRegNet / regression on band + halo, rays banded):
    every rank's band must EQUAL the rows of the single-GPU frame bit for bit -- the halo covers the regulariser's
    receptive field and the kernels' arithmetic does not depend on where a tile sits.  All ranks simulated on one GPU."""
    from enerf_b200 import config as bcfg
    from enerf_b200.dist import BandShardedRenderer
    from enerf_b200.network import Network
    cfg = bcfg.make_cfg(volume_planes=[16, 8], render_if=[False, True])
    H, W = 256, 160
    sd, batch = harness.make_case(H, W, 3, cfg, seed=4)
    net = Network()
    net.load_state_dict(sd)
    net = net.cuda().eval()
    net.precision = precision
    gb = {k: v.cuda() for k, v in batch.items() if not k.startswith("rays_")}
    with torch.no_grad():
        full = {k: v.clone() for k, v in net(gb).items()}
        for rank in range(world):
            r = BandShardedRenderer(net, 1, 2, W, H, W // 2, H // 2, rank, world, device="cuda")
            r.buf.fill_(float("nan"))
            r.render_local(gb)
            torch.cuda.synchronize()
            v = r.local_views()
            n, nv = r.n_local, r.n_vol
            assert torch.equal(v["rgb"], full["rgb_level1"][0, rank * n:(rank + 1) * n]), (rank, "rgb")
            assert torch.equal(v["depth"], full["depth_level1"][0, rank * n:(rank + 1) * n]), (rank, "depth")
            assert torch.equal(v["weights"], full["weights_level1"][0, rank * n:(rank + 1) * n]), (rank, "weights")
            assert torch.equal(v["mvs_band"].reshape(-1), full["depth_mvs_level1"].reshape(-1)[rank * nv:(rank + 1) * nv]), (rank, "mvs")
            assert torch.equal(v["std_band"].reshape(-1), full["std_level1"].reshape(-1)[rank * nv:(rank + 1) * nv]), (rank, "std")
    assert net.band_shard is False and net.ray_rows is None and net.output_views is None


def test_static_mask_is_graph_capturable_and_equal(harness):
    """network_human without the host read-back of the masked-ray count: same rgb, depth / weights equal on the
    first `mask_count` rows (zeros after), and the whole forward replays as one CUDA graph."""
    from _helpers import load_golden
    from enerf_b200 import config as bcfg
    from enerf_b200.network_human import Network
    from enerf_b200.pipeline import GraphedNetwork
    fx = load_golden("c4_human_small")
    bcfg.set_cfg(fx["cfg"])
    net = Network()
    net.load_state_dict(fx["state_dict"])
    net = net.cuda().eval()
    gb = {k: v.cuda() for k, v in fx["batch"].items()}
    with torch.no_grad():
        ref = {k: v.clone() for k, v in net(gb).items()}
        net.static_mask = True
        out = {k: v.clone() for k, v in net(gb).items()}
        g = GraphedNetwork(net, gb)
        rep = g(gb)
    torch.cuda.synchronize()
    n = int(out["mask_count"].item())
    assert n == ref["depth_level1"].shape[1] == int(fx["batch"]["mask_at_box"].bool().sum())
    for o in (out, rep):
        assert torch.equal(o["rgb_level1"], ref["rgb_level1"])
        assert torch.equal(o["depth_level1"][:, :n], ref["depth_level1"]) and not o["depth_level1"][:, n:].any()
        assert torch.equal(o["weights_level1"][:, :n], ref["weights_level1"]) and not o["weights_level1"][:, n:].any()
        assert torch.equal(o["depth_mvs_level1"], ref["depth_mvs_level1"])
    # a single masked ray is not scattered (network_human.py:104) and an empty mask renders zeros
    for keep in (1, 0):
        m = torch.zeros_like(gb["mask_at_box"])
        if keep:
            m.view(-1)[777] = 1
        b2 = dict(gb)
        b2["mask_at_box"] = m
        with torch.no_grad():
            o = net(b2)
        assert int(o["mask_count"].item()) == keep and not o["rgb_level1"].any()


def test_mask_compaction_edge_cases():
    """Order-preserving compaction == rays[mask] for empty / full / single / ragged masks and every
    mask dtype the data layer produces (bool, uint8, int32, int64)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enerf_b200 import capi
    g = torch.Generator().manual_seed(0)
    for n in (1, 31, 1024, 1025, 70001):
        rays = torch.randn(n, 8, generator=g).cuda()
        for kind in ("empty", "full", "single", "random"):
            m = {"empty": torch.zeros(n), "full": torch.ones(n), "single": torch.zeros(n), "random": (torch.rand(n, generator=g) < 0.3).float()}[kind]
            if kind == "single":
                m[n // 2] = 1
            for dt in (torch.bool, torch.uint8, torch.int32, torch.int64):
                mask = m.to(dt).cuda()
                idx = torch.full((n,), -1, dtype=torch.int32, device="cuda")
                out = torch.zeros(n, 8, device="cuda")
                cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
                ws = torch.zeros(capi.mask_compact_workspace_bytes(n) // 4 + 1, device="cuda")
                capi.mask_compact(mask, rays, idx, out, cnt, ws)
                k = int(cnt.item())
                ref = torch.nonzero(m.bool()).reshape(-1)
                assert k == ref.numel(), (n, kind, dt)
                assert torch.equal(idx[:k].cpu().long(), ref)
                assert torch.equal(out[:k].cpu(), rays.cpu()[m.bool()])
                if k:
                    dst = torch.zeros(n, 3, device="cuda")
                    src = torch.randn(k, 3, generator=g).cuda()
                    capi.scatter_rows(src, idx, k, dst)
                    exp = torch.zeros(n, 3)
                    exp[m.bool()] = src.cpu()
                    assert torch.equal(dst.cpu(), exp)


def test_device_psnr_and_rgb8_pack():
    """enerf_psnr_accumulate == the evaluator's masked PSNR; enerf_pack_rgb8 == (img*255).to(uint8) (+flip)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import math
    from enerf_b200 import capi
    g = torch.Generator().manual_seed(4)
    H, W = 48, 80
    pred, gt = torch.rand(H * W, 3, generator=g), torch.rand(H * W, 3, generator=g)
    msk = (torch.rand(H * W, generator=g) < 0.6)
    for mask in (None, msk.to(torch.uint8), msk.to(torch.int32)):
        acc = torch.zeros(2, dtype=torch.float64, device="cuda")
        capi.psnr_accumulate(pred.cuda(), gt.cuda(), mask.cuda() if mask is not None else None, acc)
        sse, cnt = acc.cpu().tolist()
        sel = msk if mask is not None else torch.ones_like(msk)
        ref_mse = ((pred[sel] - gt[sel]).double() ** 2).mean().item()
        assert cnt == sel.sum().item() * 3
        assert abs(10 * math.log10(cnt / sse) - 10 * math.log10(1.0 / ref_mse)) < 1e-6
    rgb = (torch.rand(H * W, 3, generator=g) * 1.2 - 0.1)
    for flip in (False, True):
        out = torch.zeros(H, W, 3, dtype=torch.uint8, device="cuda")
        capi.pack_rgb8(rgb.cuda(), H, W, out, flip_vertical=flip)
        ref = (rgb.clamp(0, 1) * 255).to(torch.uint8).view(H, W, 3)
        if flip:
            ref = torch.flip(ref, (0,))
        assert torch.equal(out.cpu(), ref)
