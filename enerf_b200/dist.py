"""Multi-GPU plumbing for the render path: one process per GPU (torchrun), NCCL over NVLink.

Two ways to use N GPUs (DESIGN.md section "Multi-GPU"):

* ``RayShardedRenderer`` -- the north-star layout: the target view's rays are split into N
  contiguous row bands (rays are independent), every rank renders its band
  and ONE ``all_gather`` reassembles the frame.  Each rank's kernel writes rgb|depth|weights
  straight into its own segment of the gather buffer, so there is no pack copy before the
  collective.  The per-frame front end (FeatureNet, cost volumes, 3-D CNNs) is replicated, which
  bounds the intra-frame speed-up (Amdahl; measured numbers in DESIGN.md).
* ``FrameParallelRenderer`` -- sequence rendering: rank r renders frames r, r+N, ... of a sequence
  (independent units, no data-path collective) and one ``all_gather`` per step collects the N
  finished frames.  This is the throughput mode bench.py reports at N > 1.

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
