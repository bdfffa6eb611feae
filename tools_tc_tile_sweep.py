"""Tuning sweep for the tcgen05 convolution kernel: every stride-1 3x3 layer shape of the headline frame
x {kx folding off/on} x tile candidates (enerf_tc_conv_tune), CUDA-event time per launch.
Output: one line per (layer, fold, TZ, TY) and the winners; used to set the tile rule in csrc/tc_conv.cu."""
import json
import sys

import torch

sys.path.insert(0, "/root/repo")
from enerf_b200 import capi, packing  # noqa: E402

LAYERS = [  # name, KD, cin, cout, mode, relu, (D,H,W)
    ("feat.conv0.1", 1, 8, 8, 0, 1, (3, 512, 640)),
    ("feat.conv1.1", 1, 16, 16, 0, 1, (3, 256, 320)),
    ("feat.conv2.1", 1, 32, 32, 0, 1, (3, 128, 160)),
    ("feat.smooth1", 1, 32, 16, 0, 0, (3, 256, 320)),
    ("feat.smooth0", 1, 32, 8, 0, 0, (3, 512, 640)),
    ("reg0.conv0", 3, 32, 8, 0, 1, (48, 64, 80)),
    ("reg0.conv2", 3, 16, 16, 0, 1, (24, 32, 40)),
    ("reg0.conv4", 3, 32, 32, 0, 1, (12, 16, 20)),
    ("reg0.head1", 3, 8, 1, 3, 0, (48, 64, 80)),
    ("reg1.conv0", 3, 16, 8, 0, 1, (8, 256, 320)),
    ("reg1.conv2", 3, 16, 16, 0, 1, (4, 128, 160)),
    ("reg1.conv4", 3, 32, 32, 0, 1, (2, 64, 80)),
    ("reg1.conv6", 3, 64, 64, 0, 1, (1, 32, 40)),
    ("reg1.head9", 3, 8, 9, 1, 0, (8, 256, 320)),
]
TILES_2D = [(1, 3), (1, 4), (1, 7), (1, 8), (1, 11), (1, 15), (1, 16)]
TILES_3D = [(1, 4), (1, 8), (1, 15), (2, 3), (2, 4), (2, 7), (2, 8), (3, 4), (4, 2), (4, 3), (4, 4), (4, 8), (8, 2)]


# second sweep: the strided, transposed and 1x1 layers (kind, KD, KH, stride, cin, cout, mode, relu, INPUT dims)
LAYERS2 = [
    ("feat.conv1.0 5x5s2", 0, 1, 5, 2, 8, 16, 0, 1, (3, 512, 640)),
    ("feat.conv2.0 5x5s2", 0, 1, 5, 2, 16, 32, 0, 1, (3, 256, 320)),
    ("feat.toplayer 1x1", 0, 1, 1, 1, 32, 32, 0, 0, (3, 128, 160)),
    ("reg0.conv1 s2", 0, 3, 3, 2, 8, 16, 0, 1, (48, 64, 80)),
    ("reg0.conv3 s2", 0, 3, 3, 2, 16, 32, 0, 1, (24, 32, 40)),
    ("reg1.conv1 s2", 0, 3, 3, 2, 8, 16, 0, 1, (8, 256, 320)),
    ("reg1.conv3 s2", 0, 3, 3, 2, 16, 32, 0, 1, (4, 128, 160)),
    ("reg1.conv5 s2", 0, 3, 3, 2, 32, 64, 0, 1, (2, 64, 80)),
    ("reg0.conv9 dec", 1, 3, 3, 1, 32, 16, 2, 0, (12, 16, 20)),
    ("reg0.conv11 dec", 1, 3, 3, 1, 16, 8, 2, 0, (24, 32, 40)),
    ("reg1.conv7 dec", 1, 3, 3, 1, 64, 32, 2, 0, (1, 32, 40)),
    ("reg1.conv9 dec", 1, 3, 3, 1, 32, 16, 2, 0, (2, 64, 80)),
    ("reg1.conv11 dec", 1, 3, 3, 1, 16, 8, 2, 0, (4, 128, 160)),
    ("reg1.head9", 0, 3, 3, 1, 8, 9, 1, 0, (8, 256, 320)),
]


def sweep2():
    results = []
    for name, kind, KD, KH, stride, cin, cout, mode, relu, (D, H, W) in LAYERS2:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(D, H, W, cin, generator=g).cuda()
        skip = None
        if kind == 0:
            w = torch.randn(cout, cin, KD, KH, KH, generator=g) / (cin * KD * KH * KH) ** 0.5
            wp = packing.pack_tc_conv(packing._taps_cin_cout(w)).cuda()
            Do, Ho, Wo = (D // 2 if (stride == 2 and KD > 1) else D), H // stride, W // stride
        else:
            w = torch.randn(cin, cout, 3, 3, 3, generator=g) / (cin * 27 / 8) ** 0.5
            wp = packing.pack_tc_deconv(w.permute(2, 3, 4, 0, 1).reshape(27, cin, cout)).cuda()
            Do, Ho, Wo = 2 * D, 2 * H, 2 * W
            skip = torch.randn(Do, Ho, Wo, cout, generator=g).cuda()
        bias = torch.zeros(cout).cuda() if mode in (0, 2) else None
        out = torch.empty(Do, Ho, Wo, 8 if mode == 1 else cout).cuda()
        out2 = torch.empty(Do, Ho, Wo).cuda() if mode == 1 else None
        best = None
        three_d = (KD == 3) or kind == 1
        for tz, ty in [(0, 0)] + (TILES_3D if three_d else TILES_2D):
            capi.tc_conv_tune(tz, ty, 0)
            try:
                fn = lambda: capi.tc_conv(kind, KD, KH, cout, mode, relu, x, wp, bias, skip, out, out2, out_cstride=(8 if mode == 1 else cout), stride=stride)
                us = time_launch(fn)
            except Exception:
                us = None
                torch.cuda.synchronize()
            results.append({"layer": name, "tz": tz, "ty": ty, "us": us})
            print(f"{name:20s} tile=({tz},{ty}) {'%.1f' % us if us else 'n/a'}", flush=True)
            if us and (best is None or us < best[0]):
                best = (us, tz, ty)
        print(f"== {name}: best {best}", flush=True)
    capi.tc_conv_tune(0, 0, -1)
    json.dump(results, open("/root/repo/gpurun_out/tile_sweep2.json", "w"))


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
    return e0.elapsed_time(e1) / reps * 1000.0   # us


def main():
    results = []
    for name, KD, cin, cout, mode, relu, (D, H, W) in LAYERS:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(D, H, W, cin, generator=g).cuda()
        w = torch.randn(cout, cin, KD, 3, 3, generator=g) / (cin * KD * 9) ** 0.5
        bias = torch.zeros(cout).cuda() if mode == 0 else None
        out = torch.empty(D, H, W, 8 if mode == 1 else (cout if mode == 0 else 1)).cuda()
        out2 = torch.empty(D, H, W).cuda() if mode == 1 else None
        if mode == 3:
            out = torch.empty(D, H, W).cuda()
        best = {}
        for fold in (0, 1):
            wp = packing.pack_tc_conv(packing._taps_cin_cout(w), fold_kx=bool(fold)).cuda()
            for tz, ty in [(0, 0)] + (TILES_2D if KD == 1 else TILES_3D):
                if tz > D:
                    continue
                capi.tc_conv_tune(tz, ty, fold)
                try:
                    fn = lambda: capi.tc_conv(0, KD, 3, cout, mode, relu, x, wp, bias, None, out, out2, out_cstride=(8 if mode == 1 else cout))
                    us = time_launch(fn)
                except Exception as e:  # does not fit
                    us = None
                    torch.cuda.synchronize()
                results.append({"layer": name, "fold": fold, "tz": tz, "ty": ty, "us": us})
                print(f"{name:14s} fold={fold} tile=({tz},{ty}) {'%.1f' % us if us else 'n/a'}", flush=True)
                if us and (fold not in best or us < best[fold][0]):
                    best[fold] = (us, tz, ty)
        print(f"== {name}: best unfolded {best.get(0)}, best folded {best.get(1)}", flush=True)
    capi.tc_conv_tune(0, 0, -1)
    json.dump(results, open("/root/repo/gpurun_out/tile_sweep.json", "w"))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "2":
        sweep2()
    else:
        main()
