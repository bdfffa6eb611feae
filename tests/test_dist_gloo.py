"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row-band sharding + single allgather
reassembly (RayShardedRenderer) and frame-parallel collection (FrameParallelRenderer)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from enerf_b200.dist import FrameParallelRenderer, RayShardedRenderer, row_band, segment_layout

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
