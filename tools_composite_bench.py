"""Measurement of the layered (network_composite) path at the size of the reference's enerf_outdoor
config (configs/enerf/enerf_outdoor/actor1.yaml: 768x1024 input, 3 source views, volume_planes
[32, 8], num_samples [2, 1], one foreground layer + background, both cascade levels rendered).

Prints one JSON line: frames/s of enerf_b200.network_composite.Network (eager launches and CUDA-graph
replay, device-resident inputs, CUDA-event timing, W warm-up + K timed frames) and, with
--cpu-frames N, the CPU oracle (oracle/enerf_oracle_composite.py) on the same inputs.
Synthetic data: random-init weights with randomised BN statistics, U(-1,1) images (enerf_b200/synthetic.py).
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--layers", type=int, default=1)
    ap.add_argument("--hw", default="768,1024")
    ap.add_argument("--cpu-frames", type=int, default=0)
    ap.add_argument("--precision", default="tf32")
    args = ap.parse_args()
    from enerf_b200 import capi, config as bcfg, synthetic
    from enerf_b200.network_composite import Network
    from enerf_b200.pipeline import GraphedNetwork

    H, W = (int(v) for v in args.hw.split(","))
    cfg = bcfg.set_cfg(bcfg.composite_cfg(num_fg_layers=args.layers))
    torch.manual_seed(0)
    net = Network()
    synthetic.randomize_bn_(net, seed=1)
    net = net.cuda().eval()
    net.precision = args.precision
    batch = synthetic.make_composite_batch(H, W, 3, cfg, seed=2)
    dbatch = {k: (v.cuda() if k != "bbox" else v) for k, v in batch.items() if not k.startswith("rays_")}

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    with torch.no_grad():
        n0 = capi.LAUNCHES
        out = net(dbatch)
        launches = capi.LAUNCHES - n0
        eager_ms = timed(lambda: net(dbatch))
        g = GraphedNetwork(net, dbatch)
        graph_ms = timed(g.replay)
    line = {"metric": "rendered_fps_composite", "unit": "frames/s", "value": 1000.0 / graph_ms, "eager_fps": 1000.0 / eager_ms,
            "ms_per_frame": graph_ms, "gpu_launches": launches, "dtype": "tf32" if args.precision == "tf32" else "f32",
            "config": {"workload": f"network_composite {H}x{W} S=3 planes [32,8]+bg[16,4] samples [2,1] fg_layers={args.layers} "
                                   f"bbox={batch['bbox'][0].int().tolist()}", "steps": args.steps, "warmup": args.warmup},
            "data": "synthetic"}
    if args.cpu_frames > 0:
        from oracle import enerf_oracle_composite as OC
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        torch.set_num_threads(min(32, os.cpu_count() or 1))
        with torch.no_grad():
            ref = OC.forward(sd, cfg, batch)
            t0 = time.perf_counter()
            for _ in range(args.cpu_frames):
                OC.forward(sd, cfg, batch)
            cpu_s = (time.perf_counter() - t0) / args.cpu_frames
        rgb = out["rgb_level1"].cpu()
        line["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{args.cpu_frames} frame(s) of the same workload"}
        line["parity"] = {"max_abs_rgb": (rgb - ref["rgb_level1"]).abs().max().item(),
                          "psnr_vs_oracle_db": synthetic.psnr(rgb, ref["rgb_level1"])}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
