"""Per-layer A/B of the two tcgen05 convolution kernels at the headline frame's layer shapes: csrc/tc_conv.cu (v1)
against the persistent TMA-fed csrc/tc_conv2.cu (v2) in its variants (MMA-issuing warps, CTAs per SM, tile, K-block).
CUDA-event time per launch (20 reps after 3 warm-ups), input re-randomised never (L2-warm, like inside a frame).
Writes gpurun_out/conv2_sweep.json; `python tools_conv2_sweep.py quick` runs only the default variants."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enerf_b200 import capi, packing  # noqa: E402

# name, kind, KD, KH, cin, cout, mode, relu, (D,H,W) of the ROW grid
LAYERS = [
    ("feat.conv0.1", 0, 1, 3, 8, 8, 0, 1, (3, 512, 640)),
    ("feat.conv1.1", 0, 1, 3, 16, 16, 0, 1, (3, 256, 320)),
    ("feat.conv2.1", 0, 1, 3, 32, 32, 0, 1, (3, 128, 160)),
    ("feat.toplayer", 0, 1, 1, 32, 32, 0, 0, (3, 128, 160)),
    ("feat.smooth1", 0, 1, 3, 32, 16, 0, 0, (3, 256, 320)),
    ("feat.smooth0", 0, 1, 3, 32, 8, 0, 0, (3, 512, 640)),
    ("reg0.conv0", 0, 3, 3, 32, 8, 0, 1, (48, 64, 80)),
    ("reg0.conv2", 0, 3, 3, 16, 16, 0, 1, (24, 32, 40)),
    ("reg0.conv11", 1, 3, 3, 16, 8, 2, 0, (24, 32, 40)),
    ("reg0.head1", 0, 3, 3, 8, 1, 3, 0, (48, 64, 80)),
    ("reg1.conv0", 0, 3, 3, 16, 8, 0, 1, (8, 256, 320)),
    ("reg1.conv2", 0, 3, 3, 16, 16, 0, 1, (4, 128, 160)),
    ("reg1.conv4", 0, 3, 3, 32, 32, 0, 1, (2, 64, 80)),
    ("reg1.conv9", 1, 3, 3, 32, 16, 2, 0, (2, 64, 80)),
    ("reg1.conv11", 1, 3, 3, 16, 8, 2, 0, (4, 128, 160)),
    ("reg1.head9", 0, 3, 3, 8, 9, 1, 0, (8, 256, 320)),
]


def time_launch(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main(quick):
    variants = [("v1", dict(impl=1)), ("v2", dict(impl=0, nmma=1, ctas_per_sm=1)), ("v2.nmma2", dict(impl=0, nmma=2, ctas_per_sm=1)),
                ("v2.2cta", dict(impl=0, ctas_per_sm=2, nmma=1)), ("v2.2cta.nmma2", dict(impl=0, ctas_per_sm=2, nmma=2))]
    if not quick:
        for kbc in (8, 16, 32):
            variants.append((f"v2.2cta.nmma2.kbc{kbc}", dict(impl=0, kbc=kbc)))
        for tz, ty in ((1, 7), (1, 11), (1, 15), (2, 4), (2, 8), (4, 4), (4, 8), (2, 3)):
            variants.append((f"v2.2cta.nmma2.t{tz}x{ty}", dict(impl=0, tz=tz, ty=ty)))
            variants.append((f"v2.1cta.nmma2.t{tz}x{ty}", dict(impl=0, ctas_per_sm=1, tz=tz, ty=ty)))
    results, totals = [], {}
    for name, kind, KD, KH, cin, cout, mode, relu, (D, H, W) in LAYERS:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(D, H, W, cin, generator=g).cuda()
        skip = None
        if kind == 0:
            w = torch.randn(cout, cin, KD, KH, KH, generator=g) / (cin * KD * KH * KH) ** 0.5
            wp = packing.pack_tc_conv(packing._taps_cin_cout(w), fold_kx=packing.tc_fold_kx(KD, KH, 1, cout, single=(mode == 3), head=(mode == 1), cin=cin)).cuda()
            Do, Ho, Wo = D, H, W
        else:
            w = torch.randn(cin, cout, 3, 3, 3, generator=g) / (cin * 27 / 8) ** 0.5
            wp = packing.pack_tc_deconv(w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout)).cuda()
            Do, Ho, Wo = 2 * D, 2 * H, 2 * W
            skip = torch.randn(Do, Ho, Wo, cout, generator=g).cuda()
        bias = torch.zeros(cout).cuda() if mode in (0, 2) else None
        out = torch.empty((Do, Ho, Wo) if mode == 3 else (Do, Ho, Wo, 8 if mode == 1 else cout)).cuda()
        out2 = torch.empty(Do, Ho, Wo).cuda() if mode == 1 else None
        ref = None
        layer_variants = list(variants)
        rule2 = kind == 0 and KD == 1 and KH == 3 and cout == 8      # layers the level-2 rule folds
        rule1 = kind == 0 and KD == 3 and mode == 1                  # ... and the level-1 rule (feat + prob head)
        if rule2 or rule1:
            layer_variants += [("v2.2cta.rule0", dict(impl=0, _rule=0)), ("v2.1cta.rule0", dict(impl=0, ctas_per_sm=1, _rule=0)), ("v1.rule0", dict(impl=1, _rule=0)),
                               ("v2.1cta", dict(impl=0, ctas_per_sm=1))]
            for tz, ty in (((1, 7), (1, 11), (1, 15)) if rule2 else ((2, 4), (2, 8), (4, 4))):
                layer_variants.append((f"v2.2cta.t{tz}x{ty}", dict(impl=0, tz=tz, ty=ty)))
        if kind == 0 and KD == 3 and packing.tc_fold_kx(KD, KH, 1, cout, single=(mode == 3)):       # the always-folded 3-D layers, unfolded (27 taps of N = C)
            layer_variants += [("v2.2cta.nmma2.nofold", dict(impl=0, _nofold=True)), ("v1.nofold", dict(impl=1, _nofold=True))]
        for vname, kw in layer_variants:
            is3d = KD == 3 or kind == 1
            if ".t1x" in vname and is3d or (".t2x" in vname or ".t4x" in vname) and not is3d:
                continue
            kw = dict(kw)
            rule = kw.pop("_rule", None)
            nofold = kw.pop("_nofold", False)
            capi.tc_conv2_tune(**kw)
            wp_v = wp
            if rule is not None:
                capi.tc_conv_fold_rule(rule)
                wp_v = packing.pack_tc_conv(packing._taps_cin_cout(w), fold_kx=packing.tc_fold_kx(KD, KH, 1, cout, single=(mode == 3), head=(mode == 1), cin=cin)).cuda()
            if nofold:
                wp_v = packing.pack_tc_conv(packing._taps_cin_cout(w), fold_kx=False).cuda()
                capi.tc_conv_tune(0, 0, 0)
            try:
                out.fill_(float("nan"))
                fn = lambda: capi.tc_conv(kind, KD, KH, cout, mode, relu, x, wp_v, bias, skip, out, out2, out_cstride=(8 if mode == 1 else cout))  # noqa: E731
                us = time_launch(fn)
                if ref is None:
                    ref = out.clone()
                    same = True
                else:
                    same = bool(torch.equal(out, ref)) if not (rule is not None or nofold) else float((out - ref).abs().max())
            except Exception as e:  # noqa: BLE001
                us, same = None, str(e)[:120]
                torch.cuda.synchronize()
            finally:
                if nofold:
                    capi.tc_conv_tune(0, 0, -1)
                if rule is not None:
                    capi.tc_conv_fold_rule(2)
            results.append({"layer": name, "variant": vname, "us": us, "equal_to_v1": same})
            if us and vname in ("v1", "v2", "v2.nmma2", "v2.2cta", "v2.2cta.nmma2"):
                totals[vname] = totals.get(vname, 0.0) + us
            print(f"{name:16s} {vname:18s} {'%.1f us' % us if us else 'n/a':>10s}  equal={same}", flush=True)
    capi.tc_conv2_tune()
    print("totals (us):", {k: round(v, 1) for k, v in totals.items()})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"results": results, "totals_us": totals}, open("gpurun_out/conv2_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main(quick=(len(sys.argv) > 1 and sys.argv[1] == "quick"))
