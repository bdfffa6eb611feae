"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row-band sharding + single allgather
reassembly (RayShardedRenderer) and frame-parallel collection (FrameParallelRenderer)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from enerf_b200.dist import (BandShardedRenderer, FrameParallelRenderer, RayShardedRenderer, amdahl_bound, band_segment_layout, row_band,
                             segment_layout)

W, H, NS = 12, 8, 2


def _fake_render(batch, out, level=1):
    r = batch[f"rays_{level}"][0]                      # (n, 8): a deterministic function of each ray
    out["rgb"].copy_(torch.stack([r[:, 6] * 0.01, r[:, 7] * 0.02, r[:, 6] + r[:, 7]], -1))
    out["depth"].copy_(r[:, 6] * 100 + r[:, 7])
    out["weights"].copy_(torch.stack([r[:, 6], -r[:, 7]], -1))


def _full_rays():
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    rays = torch.zeros(H * W, 8)
    rays[:, 6], rays[:, 7] = xs.reshape(-1), ys.reshape(-1)
    return rays[None]


class _FakeNet:
    """Stands in for enerf_b200.network.Network on CPU: honours ray_rows / output_views / band_shard the way the
    real forward does (its band is a deterministic function of the pixel coordinates)."""
    output_views, ray_rows, band_shard = None, None, False

    def __call__(self, batch):
        r0, r1 = self.ray_rows
        ys, xs = torch.meshgrid(torch.arange(r0, r1, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        rays = torch.zeros((r1 - r0) * W, 8)
        rays[:, 6], rays[:, 7] = xs.reshape(-1), ys.reshape(-1)
        _fake_render({"rays_1": rays[None]}, self.output_views[1])
        hv, wv = H // 2, W // 2
        vy, vx = torch.meshgrid(torch.arange(hv, dtype=torch.float32), torch.arange(wv, dtype=torch.float32), indexing="ij")
        mvs, std = vy * 10 + vx, vy - vx
        if self.band_shard:          # rows outside the band are not this rank's: poison them
            v0, v1 = r0 // 2, r1 // 2
            mvs[:v0], mvs[v1:], std[:v0], std[v1:] = -1.0, -1.0, -1.0, -1.0
        return {"depth_mvs_level1": mvs[None], "std_level1": std[None]}


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = {"rays_1": _full_rays()}
        ref = {"rgb": torch.empty(H * W, 3), "depth": torch.empty(H * W), "weights": torch.empty(H * W, NS)}
        _fake_render(batch, ref)
        r = RayShardedRenderer(_fake_render, 1, NS, W, H, rank, world)
        out = r(batch)
        assert torch.equal(out["rgb"][0], ref["rgb"]) and torch.equal(out["depth"][0], ref["depth"])
        assert torch.equal(out["weights"][0], ref["weights"])
        # band layout: ray outputs + the band rows of depth_mvs / std travel in ONE all-gather
        net = _FakeNet()
        b = BandShardedRenderer(net, 1, NS, W, H, W // 2, H // 2, rank, world)
        got = b.render({"src_inps": torch.zeros(1)})
        assert net.ray_rows is None and net.output_views is None and net.band_shard is False
        assert torch.equal(got["rgb_level1"][0], ref["rgb"]) and torch.equal(got["depth_level1"][0], ref["depth"])
        assert torch.equal(got["weights_level1"][0], ref["weights"])
        net1 = _FakeNet()
        net1.ray_rows, net1.output_views = (0, H), {1: {k: torch.empty_like(v) for k, v in ref.items()}}
        whole = net1({})
        assert torch.equal(got["depth_mvs_level1"], whole["depth_mvs_level1"]) and torch.equal(got["std_level1"], whole["std_level1"])
        # frame parallel: rank r renders frame r (rays offset by r), everyone receives both frames
        fb = {"rays_1": _full_rays() + rank}
        f = FrameParallelRenderer(_fake_render, H * W, NS, rank, world)
        frames = f(fb)
        for q in range(world):
            exp = {"rgb": torch.empty(H * W, 3), "depth": torch.empty(H * W), "weights": torch.empty(H * W, NS)}
            _fake_render({"rays_1": _full_rays() + q}, exp)
            assert torch.equal(frames["rgb"][q], exp["rgb"]) and torch.equal(frames["depth"][q], exp["depth"])
    finally:
        dist.destroy_process_group()


def test_row_band_partition():
    for n, w in ((512, 8), (10, 3), (7, 7), (64, 1)):
        bands = [row_band(n, r, w) for r in range(w)]
        assert bands[0][0] == 0 and bands[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        sizes = [b - a for a, b in bands]
        assert max(sizes) - min(sizes) <= 1
    assert segment_layout(10, 2) == (0, 30, 40, 60)
    assert band_segment_layout(10, 2, 4) == (0, 30, 40, 60, 64, 68)
    st = {"feature_net": 0.4, "cost_volume_0": 0.05, "cost_reg_0": 0.2, "cost_volume_1": 0.05, "cost_reg_1": 0.35, "render_rays_1": 0.3}
    assert abs(amdahl_bound(st, 1, 1.0) - 1.0) < 1e-12
    assert abs(amdahl_bound(st, 8, 0.375) - 1.35 / (0.65 + 0.4 * 0.375 + 0.3 / 8)) < 1e-12


def test_ray_sharding_and_frame_parallel_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)


def test_single_process_identity():
    batch = {"rays_1": _full_rays()}
    out = RayShardedRenderer(_fake_render, 1, NS, W, H)(batch)
    ref = {"rgb": torch.empty(H * W, 3), "depth": torch.empty(H * W), "weights": torch.empty(H * W, NS)}
    _fake_render(batch, ref)
    assert torch.equal(out["rgb"][0], ref["rgb"])
