"""Multi-GPU plumbing for the render path: one process per GPU (torchrun), NCCL over NVLink.

Two ways to use N GPUs (DESIGN.md section "Multi-GPU"):

* ``RayShardedRenderer`` -- the north-star layout: the target view's rays are split into N
  contiguous row bands (rays are independent), every rank renders its band
  and ONE ``all_gather`` reassembles the frame.  Each rank's kernel writes rgb|depth|weights
  straight into its own segment of the gather buffer, so there is no pack copy before the
  collective.  The per-frame front end (FeatureNet, cost volumes, 3-D CNNs) is replicated, which
  bounds the intra-frame speed-up (Amdahl; measured numbers in DESIGN.md).
* ``BandShardedRenderer`` -- the same layout with the per-frame front end sharded as far as it can be
  without approximation (SURVEY.md section 8e option 1): the level-1 cost volume, CostRegNet and depth
  regression run only on the rank's rows plus a halo covering the regulariser's receptive field, so each
  band is BIT-IDENTICAL to the single-GPU frame; FeatureNet and the coarse level stay replicated.  One
  all-gather reassembles rgb | depth | weights | depth_mvs | std.
* ``FrameParallelRenderer`` -- sequence rendering: rank r renders frames r, r+N, ... of a sequence
  (independent units, no data-path collective; every rank hands its own frames to its consumer).
  This is the throughput mode bench.py reports at N > 1.  ``gather()`` is optional (one collective
  per step when a single consumer needs every frame).

The reference has no multi-GPU inference at all (run.py:23,48 put the model on one device).
The host logic is backend-agnostic and is covered on CPU with gloo, world_size 2
(tests/test_dist_gloo.py).
"""
import os

import torch
import torch.distributed as dist


def init_from_env():
    """torchrun / torch.distributed.run environment -> (rank, local_rank, world).  No-op for world 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        use_cuda = torch.cuda.is_available()
        if use_cuda:
            torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl" if use_cuda else "gloo", init_method="env://")
    return rank, local, world


def row_band(n_rows, rank, world):
    """Contiguous, balanced split of ``n_rows`` image rows: rows [r0, r1) for ``rank``."""
    base, rem = divmod(n_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def segment_layout(n_rays, n_samples):
    """Float offsets of rgb | depth | weights inside one rank's gather segment."""
    o_rgb, o_depth, o_w = 0, 3 * n_rays, 4 * n_rays
    return o_rgb, o_depth, o_w, (4 + n_samples) * n_rays


class RayShardedRenderer:
    """render_fn(batch, out) -> None renders ``batch['rays_<level>']`` into the views in ``out``
    (keys rgb (n,3), depth (n,), weights (n,Ns)); everything else about the frame is replicated."""

    def __init__(self, render_fn, level, n_samples, width, height, rank=0, world=1, group=None, device="cpu"):
        self.render_fn, self.level, self.ns = render_fn, level, n_samples
        self.W, self.H, self.rank, self.world, self.group = width, height, rank, world, group
        if height % world:
            raise ValueError(f"ray sharding needs the render height {height} divisible by the world size {world}")
        self.rows = height // world
        self.n_local = self.rows * width
        *_, self.seg = segment_layout(self.n_local, n_samples)
        self.buf = torch.empty(world * self.seg, device=device, dtype=torch.float32)

    def local_views(self):
        o_rgb, o_depth, o_w, seg = segment_layout(self.n_local, self.ns)
        mine = self.buf[self.rank * seg:(self.rank + 1) * seg]
        return {"rgb": mine[o_rgb:o_depth].view(self.n_local, 3), "depth": mine[o_depth:o_w],
                "weights": mine[o_w:seg].view(self.n_local, self.ns)}

    def local_batch(self, batch):
        """This rank's share of the frame: the row band of ``rays_<level>`` (or, when the batch carries
        no rays, the band is generated on device -- see ``rows``)."""
        r0, r1 = self.rows_range()
        key = f"rays_{self.level}"
        local = dict(batch)
        if key in batch:
            local[key] = batch[key][:, r0 * self.W:r1 * self.W].contiguous()
        return local

    def rows_range(self):
        return row_band(self.H, self.rank, self.world)

    def gather(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.buf, self.buf[self.rank * self.seg:(self.rank + 1) * self.seg], group=self.group)
        return self.assemble()

    def __call__(self, batch):
        self.render_fn(self.local_batch(batch), self.local_views())
        return self.gather()

    def assemble(self):
        o_rgb, o_depth, o_w, seg = segment_layout(self.n_local, self.ns)
        per = self.buf.view(self.world, seg)
        return {"rgb": per[:, o_rgb:o_depth].reshape(1, -1, 3), "depth": per[:, o_depth:o_w].reshape(1, -1),
                "weights": per[:, o_w:seg].reshape(1, -1, self.ns)}


def band_segment_layout(n_rays, n_samples, n_vol):
    """Float offsets of rgb | depth | weights | depth_mvs | std inside one rank's gather segment
    (n_vol = volume pixels of the rank's band)."""
    o_rgb, o_depth, o_w = 0, 3 * n_rays, 4 * n_rays
    o_mvs = (4 + n_samples) * n_rays
    o_std = o_mvs + n_vol
    return o_rgb, o_depth, o_w, o_mvs, o_std, o_std + n_vol


class BandShardedRenderer:
    """Row-band sharding of ONE frame over ``world`` ranks with a single all-gather.

    ``net`` is the drop-in Network (enerf_b200/network.py); the batch must carry no ``rays_<level>`` (the band's
    rays are generated on device).  ``render(batch)`` = this rank's part + the collective; ``assemble()`` returns
    the reference's output dict for the full frame.  ``world = 1`` degenerates to the plain forward."""

    def __init__(self, net, level, n_samples, width, height, vol_w, vol_h, rank=0, world=1, group=None, device="cpu"):
        if height % world or vol_h % world:
            raise ValueError(f"band sharding needs {height} render rows and {vol_h} volume rows divisible by the world size {world}")
        self.net, self.level, self.ns, self.W, self.H, self.wv, self.hv = net, level, n_samples, width, height, vol_w, vol_h
        self.rank, self.world, self.group = rank, world, group
        self.rows, self.vrows = height // world, vol_h // world
        self.n_local, self.n_vol = self.rows * width, self.vrows * vol_w
        *_, self.seg = band_segment_layout(self.n_local, n_samples, self.n_vol)
        self.buf = torch.empty(world * self.seg, device=device, dtype=torch.float32)
        self.full = {}

    def rows_range(self):
        return row_band(self.H, self.rank, self.world)

    def local_views(self):
        o_rgb, o_depth, o_w, o_mvs, o_std, seg = band_segment_layout(self.n_local, self.ns, self.n_vol)
        mine = self.buf[self.rank * seg:(self.rank + 1) * seg]
        return {"rgb": mine[o_rgb:o_depth].view(self.n_local, 3), "depth": mine[o_depth:o_w], "weights": mine[o_w:o_mvs].view(self.n_local, self.ns),
                "mvs_band": mine[o_mvs:o_std].view(self.vrows, self.wv), "std_band": mine[o_std:seg].view(self.vrows, self.wv)}

    def render_local(self, batch):
        """This rank's band: the Network writes its ray outputs straight into the gather segment; the band rows of
        depth_mvs / std (computed at full height, only these rows are this rank's) are copied next to them."""
        net, v = self.net, self.local_views()
        saved = (net.output_views, net.ray_rows, net.band_shard)
        net.output_views = {self.level: {k: v[k] for k in ("rgb", "depth", "weights")}}
        net.ray_rows, net.band_shard = self.rows_range(), self.world > 1
        try:
            out = net(batch)
        finally:
            net.output_views, net.ray_rows, net.band_shard = saved
        v0 = self.rank * self.vrows
        v["mvs_band"].copy_(out[f"depth_mvs_level{self.level}"][0, v0:v0 + self.vrows])
        v["std_band"].copy_(out[f"std_level{self.level}"][0, v0:v0 + self.vrows])
        return out

    def gather(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.buf, self.buf[self.rank * self.seg:(self.rank + 1) * self.seg], group=self.group)
        return self.assemble()

    def render(self, batch):
        self.render_local(batch)
        return self.gather()

    def assemble(self):
        o_rgb, o_depth, o_w, o_mvs, o_std, seg = band_segment_layout(self.n_local, self.ns, self.n_vol)
        per = self.buf.view(self.world, seg)
        lv = self.level
        return {f"rgb_level{lv}": per[:, o_rgb:o_depth].reshape(1, -1, 3), f"depth_level{lv}": per[:, o_depth:o_w].reshape(1, -1),
                f"weights_level{lv}": per[:, o_w:o_mvs].reshape(1, -1, self.ns),
                f"depth_mvs_level{lv}": per[:, o_mvs:o_std].reshape(1, self.hv, self.wv), f"std_level{lv}": per[:, o_std:seg].reshape(1, self.hv, self.wv)}


def amdahl_bound(stage_ms, world, band_fraction):
    """Speed-up ceiling of the band layout from single-GPU stage times: FeatureNet and level 0 replicated, the level-1
    volume stages on ``band_fraction`` of the rows, the ray stage on 1/world of the rays, free collective."""
    rep = sum(v for k, v in stage_ms.items() if k.endswith("_0") or k in ("feature_net", "camera_setup", "depth_hypotheses_1", "pack_img_feat_1"))
    vol1 = stage_ms.get("cost_volume_1", 0.0) + stage_ms.get("cost_reg_1", 0.0) + stage_ms.get("depth_regress_1", 0.0)
    rays = stage_ms.get("render_rays_1", 0.0)
    total = rep + vol1 + rays
    return total / (rep + vol1 * band_fraction + rays / world) if total > 0 else None


def measure_intra_frame(net, batch, rank, world, device, steps=20, warmup=3, flush_buf=None, single_frame_ms=None):
    """Latency of ONE c2-style frame rendered by all ranks together (band layout + one all-gather), max over ranks.
    Every rank must pass the SAME frame.  Returns the dict bench.py prints as config.intra_frame."""
    from .config import get_cfg, snapshot
    from .pipeline import GraphedNetwork
    levels = snapshot(get_cfg())
    lv = levels[-1]
    H, W = batch["src_inps"].shape[-2:]
    Hr, Wr, hv, wv = int(H * lv.render_scale), int(W * lv.render_scale), int(H * lv.volume_scale), int(W * lv.volume_scale)
    batch = {k: v for k, v in batch.items() if not k.startswith("rays_")}
    r = BandShardedRenderer(net, len(levels) - 1, lv.num_samples, Wr, Hr, wv, hv, rank, world, device=device)

    g = GraphedNetwork(net, batch, fn=r.render_local)      # the graph holds the rank's band incl. the two band copies

    def frame():
        g.replay()
        r.gather()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, warmup)):
        frame()
    sync()
    ev = []
    for _ in range(steps):
        if flush_buf is not None:
            flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        frame()
        b.record()
        ev.append((a, b))
    sync()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    # Amdahl ceiling from this rank's single-GPU stage profile
    net.profile = True
    with torch.no_grad():
        net(batch)
    torch.cuda.synchronize()
    stages = net.stage_times_ms()
    net.profile = False
    deep_halo = net.HALO[True]
    band_fraction = min(1.0, (hv // world + 2 * deep_halo) / hv) if world > 1 else 1.0
    bound = amdahl_bound(stages, world, band_fraction)
    out = {"layout": f"row bands x{world}: level-1 cost volume / CostRegNet / regression on band + {deep_halo}-row halo "
                     f"({band_fraction:.3f} of the rows per rank), rays banded, FeatureNet + level 0 replicated, one NCCL all-gather "
                     f"of rgb|depth|weights|depth_mvs|std ({r.seg * 4 * world / 1e6:.1f} MB total)",
           "latency_ms": ms, "fps": 1000.0 / ms, "amdahl_bound": bound, "scaling": "strong",
           "bit_identical_to_single_gpu": "asserted by tests/test_parity_gpu.py::test_band_sharding_is_bit_identical"}
    if single_frame_ms:
        out["single_gpu_latency_ms"] = single_frame_ms
        out["speedup"] = single_frame_ms / ms
    return out


class FrameParallelRenderer:
    """Each rank renders its own frame of a sequence; one all_gather per step collects the frames."""

    def __init__(self, render_fn, n_rays, n_samples, rank=0, world=1, group=None, device="cpu"):
        self.render_fn, self.n, self.ns, self.rank, self.world, self.group = render_fn, n_rays, n_samples, rank, world, group
        *_, self.seg = segment_layout(n_rays, n_samples)
        self.buf = torch.empty(world * self.seg, device=device, dtype=torch.float32)

    def local_views(self):
        o_rgb, o_depth, o_w, seg = segment_layout(self.n, self.ns)
        mine = self.buf[self.rank * seg:(self.rank + 1) * seg]
        return {"rgb": mine[o_rgb:o_depth].view(self.n, 3), "depth": mine[o_depth:o_w], "weights": mine[o_w:seg].view(self.n, self.ns)}

    def gather(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.buf, self.buf[self.rank * self.seg:(self.rank + 1) * self.seg], group=self.group)
        o_rgb, o_depth, o_w, seg = segment_layout(self.n, self.ns)
        per = self.buf.view(self.world, seg)
        return {"rgb": per[:, o_rgb:o_depth].reshape(self.world, self.n, 3), "depth": per[:, o_depth:o_w],
                "weights": per[:, o_w:seg].reshape(self.world, self.n, self.ns)}

    def __call__(self, batch):
        self.render_fn(batch, self.local_views())
        return self.gather()


def network_render_fn(net, level, rows=None):
    """Adapter: run the drop-in Network with its ray-stage outputs redirected into ``out``
    (``rows`` = (r0, r1): the band to generate rays for when the batch carries none)."""

    def fn(batch, out):
        net.output_views, net.ray_rows = {level: out}, rows
        try:
            return net(batch)
        finally:
            net.output_views, net.ray_rows = None, None
    return fn
