"""A/B of the kernel choices at the headline frame (512x640, S=3, 48+8): per-stage CUDA-event times (eager, profiled),
single-frame latency (one CUDA graph) and the 4-frames-in-flight throughput, for combinations of
  conv: csrc/tc_conv.cu (v1) | csrc/tc_conv2.cu, lat0 fused or not, kx-fold rule level 0 / 1 / 2
  rays: csrc/render_rays_tc.cuh (single role) | csrc/render_rays_ws.cu (warp specialised)
Writes gpurun_out/frame_ab.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
from enerf_b200 import capi  # noqa: E402
from enerf_b200.pipeline import GraphedNetwork  # noqa: E402

CONFIGS = [
    # name, tc_conv2_tune kwargs, lat0 fused (False | producer warps), ray kernel, fold rule level
    ("conv v1 | rays v1 | fold rule 0 (round-1 kernels)", dict(impl=1), False, 1, 0),
    ("conv v2, lat0 fused (6 producer warps), fold rule 2 | rays v1 (shipped default)", dict(impl=0), 6, 1, 2),
    ("conv v2, lat0 fused (8 producer warps), fold rule 2 | rays v1", dict(impl=0), 8, 1, 2),
    ("conv v2, lat0 fused (4 producer warps), fold rule 2 | rays v1", dict(impl=0), 4, 1, 2),
    ("conv v2, lat0 fused (6 producer warps), fold rule 1 | rays v1", dict(impl=0), 6, 1, 1),
    ("conv v2, lat0 fused (6 producer warps), fold rule 0 | rays v1", dict(impl=0), 6, 1, 0),
    ("conv v2, lat0 separate, fold rule 2 | rays v1", dict(impl=0), False, 1, 2),
]


def main():
    dev = torch.device("cuda")
    cfg, net, batch, wl = bench.build_problem("c2")
    net = net.to(dev)
    for k in [k for k in batch if k.startswith("rays_")]:
        batch.pop(k)
    gb = bench.to_dev(batch, dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    out = []
    ref_rgb = None
    for name, conv_kw, fuse, ray_impl, rule in CONFIGS:
        capi.tc_conv2_tune(**conv_kw)
        capi.tc_conv_fold_rule(rule)
        net.invalidate_packed()
        capi.tc_conv2_fuse_lateral(fuse)
        capi.render_rays_tc_select(ray_impl)
        with torch.no_grad():
            o = net(gb)
            torch.cuda.synchronize()
            rgb = o["rgb_level1"].clone()
            if ref_rgb is None:
                ref_rgb = rgb
            net.profile = True
            acc = {}
            for _ in range(10):
                flush.zero_()
                net(gb)
                torch.cuda.synchronize()
                for k, v in net.stage_times_ms().items():
                    acc[k] = acc.get(k, 0.0) + v / 10
            net.profile = False
            reps = []
            main_stream = torch.cuda.current_stream()
            for j in range(4):
                st = torch.cuda.Stream() if j else main_stream
                with torch.cuda.stream(st):
                    reps.append((st, GraphedNetwork(net, gb), torch.cuda.Event()))
            torch.cuda.synchronize()

            def one():
                reps[0][1].replay()

            def four():
                fork = torch.cuda.Event()
                fork.record(main_stream)
                for st, g, done in reps:
                    if st is not main_stream:
                        st.wait_event(fork)
                    with torch.cuda.stream(st):
                        g.replay()
                        if st is not main_stream:
                            done.record(st)
                for st, g, done in reps[1:]:
                    main_stream.wait_event(done)

            res = {}
            for label, fn, frames in (("single_ms", one, 1), ("inflight4_ms_per_frame", four, 4)):
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
                ev = bench.timed_events(fn, 30, flush)
                torch.cuda.synchronize()
                res[label] = sum(a.elapsed_time(b) for a, b in ev) / len(ev) / frames
            del reps
        rec = {"config": name, "single_fps": 1000.0 / res["single_ms"], "inflight4_fps": 1000.0 / res["inflight4_ms_per_frame"], **res,
               "max_abs_rgb_vs_first": float((rgb - ref_rgb).abs().max()),
               "stages_ms": {k: round(v, 4) for k, v in acc.items() if v > 0.02}}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    capi.tc_conv2_tune()
    capi.tc_conv_fold_rule(2)
    capi.tc_conv2_fuse_lateral(True)
    capi.render_rays_tc_select(0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/frame_ab.json", "w"), indent=1)


if __name__ == "__main__":
    main()
