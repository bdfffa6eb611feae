"""Serving-side wrappers around the drop-in ``Network`` (no reference counterpart: the reference's
``run.py:57-76`` loop is synchronous -- H2D, forward, sync, D2H per frame).

* ``GraphedNetwork``  -- captures one ``Network.forward`` for a fixed batch signature into a CUDA
  graph (static input/output buffers) and replays it: one ``cudaGraphLaunch`` instead of ~40 kernel
  launches + the Python between them.
* ``StreamedRenderer`` -- sequence rendering from HOST batches: the H2D copy, the forward and the D2H
  copy of adjacent frames run on separate streams (and, with graphs, the forwards of ``depth`` frames
  overlap each other), so the end-to-end rate is bounded by the slowest resource instead of the sum.
  Every frame still pays its own copies.

Both produce exactly what ``Network.forward`` produces (same kernels, same order).
"""
import torch


# launch geometry the Networks read on the HOST (network_composite: one bbox per foreground layer); such
# keys stay CPU tensors, are not copied per frame and must not change under a captured graph
HOST_KEYS = ("bbox",)


def _signature(batch):
    return tuple(sorted((k, tuple(v.shape)) for k, v in batch.items() if torch.is_tensor(v)))


def _is_dev_input(k, v):
    return torch.is_tensor(v) and k not in HOST_KEYS


class GraphedNetwork:
    def __init__(self, net, example_batch, warmup=3, flat_outputs=False, fn=None):
        """example_batch: CUDA tensors; defines the static signature.
        flat_outputs: every output tensor of the captured forward lives in ONE device buffer
        (``self.flat_out``), so a consumer can fetch the whole result with a single copy.
        fn: what to capture instead of ``net(batch)`` (a callable of the static batch that drives ``net``,
        e.g. BandShardedRenderer.render_local); ``net`` still provides the scratch cache to privatise."""
        self.net = net
        call = fn if fn is not None else net
        self.sig = _signature(example_batch)
        self.static_in = {k: (v.clone() if _is_dev_input(k, v) else v) for k, v in example_batch.items() if torch.is_tensor(v)}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):           # builds packed weights / scratch, sets kernel attributes
                probe = call(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.flat_out, self.out_layout, saved_views = None, None, getattr(net, "output_views", None)
        if flat_outputs and hasattr(net, "output_views") and net.output_views is None and not getattr(net, "masked", False):
            self.out_layout, off = {}, 0
            for k, v in probe.items():
                self.out_layout[k] = (off, tuple(v.shape))
                off += (v.numel() + 63) // 64 * 64
            self.flat_out = torch.empty(off, device=next(iter(probe.values())).device, dtype=torch.float32)
            views = {}
            for k, (o, shp) in self.out_layout.items():
                name, lvl = k.rsplit("_level", 1)
                views.setdefault(int(lvl), {})[name] = self.flat_out[o:o + int(torch.tensor(shp).prod())].view(shp[1:])
            net.output_views = views
        # The captured kernels keep the ADDRESSES of the Network's scratch (camera struct, workspaces,
        # variance / probability volumes, packed image features).  Graphs of the same net replayed
        # concurrently on different streams (StreamedRenderer, bench.py replicas) must not share them:
        # the capture runs with a fresh scratch cache that this graph then owns.
        shared_scratch = getattr(net, "_buffers_cache", None)
        if shared_scratch is not None:
            net._buffers_cache = {}
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph), torch.no_grad():
                self.static_out = call(self.static_in)
        finally:
            if hasattr(net, "output_views"):
                net.output_views = saved_views
            if shared_scratch is not None:
                self.scratch, net._buffers_cache = net._buffers_cache, shared_scratch

    def load(self, batch, non_blocking=True):
        if _signature(batch) != self.sig:
            raise ValueError("batch signature differs from the captured graph")
        for k, dst in self.static_in.items():
            if k in HOST_KEYS:
                if not torch.equal(dst.cpu(), batch[k].cpu()):
                    raise ValueError(f"{k} is launch geometry baked into the captured graph; re-capture for a new value")
                continue
            dst.copy_(batch[k], non_blocking=non_blocking)

    def replay(self):
        self.graph.replay()
        return self.static_out

    def __call__(self, batch):
        self.load(batch)
        return self.replay()


class StreamedRenderer:
    """render(host_batches) -> list of dicts of pinned host tensors, one per input batch."""

    def __init__(self, net, example_host_batch, device, depth=3, use_graph=True):
        """depth = frames in flight.  With CUDA graphs every slot replays on its OWN compute stream, so
        the forwards of adjacent frames overlap as well (the coarse layers of one frame launch fewer CTAs
        than the GPU has SMs); without graphs the slots share one compute stream (Network keeps per-call
        scratch buffers)."""
        self.dev, self.depth = device, depth
        self.copy_in, self.copy_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
        shared = torch.cuda.Stream(device)
        ex = {k: (v.to(device) if _is_dev_input(k, v) else v) for k, v in example_host_batch.items() if torch.is_tensor(v)}
        self.slots = []
        for _ in range(depth):
            cs = torch.cuda.Stream(device) if use_graph else shared
            with torch.cuda.stream(cs):
                g = GraphedNetwork(net, ex, flat_outputs=True) if use_graph else None
            slot = {"g": g, "cs": cs, "in": g.static_in if g else {k: (v.clone() if _is_dev_input(k, v) else v) for k, v in ex.items()},
                    "h2d": torch.cuda.Event(), "done": torch.cuda.Event(), "d2h": torch.cuda.Event(), "free": torch.cuda.Event(),
                    "host_out": None, "host_flat": None}
            if g is not None and g.flat_out is not None:   # one pinned buffer mirrors the flat device buffer
                slot["host_flat"] = torch.empty(g.flat_out.numel(), dtype=torch.float32).pin_memory()
                slot["host_out"] = {k: slot["host_flat"][o:o + int(torch.tensor(shp).prod())].view(shp) for k, (o, shp) in g.out_layout.items()}
            self.slots.append(slot)
        self.net = net
        torch.cuda.synchronize(device)

    def _submit(self, slot, host_batch):
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(slot["free"])          # previous forward of this slot finished reading the inputs
            for k, dst in slot["in"].items():
                if k not in HOST_KEYS:
                    dst.copy_(host_batch[k], non_blocking=True)
            slot["h2d"].record(self.copy_in)
        cs = slot["cs"]
        with torch.cuda.stream(cs), torch.no_grad():
            cs.wait_event(slot["h2d"])
            cs.wait_event(slot["d2h"])                     # previous outputs of this slot have been copied out
            out = slot["g"].replay() if slot["g"] else self.net(slot["in"])
            slot["done"].record(cs)
            slot["free"].record(cs)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(slot["done"])
            if slot["host_flat"] is not None:              # the whole result in one D2H copy
                slot["host_flat"].copy_(slot["g"].flat_out, non_blocking=True)
            else:
                if slot["host_out"] is None or any(slot["host_out"][k].shape != v.shape for k, v in out.items() if torch.is_tensor(v)):
                    slot["host_out"] = {k: (torch.empty(v.shape, dtype=v.dtype).pin_memory() if torch.is_tensor(v) else v) for k, v in out.items()}
                for k, v in out.items():
                    if torch.is_tensor(v):
                        v.record_stream(self.copy_out)
                        slot["host_out"][k].copy_(v, non_blocking=True)
            slot["d2h"].record(self.copy_out)

    def render(self, host_batches, on_frame=None):
        """Streams the batches through; ``on_frame(i, host_out)`` is called once frame i is on the host
        (the pinned buffers are reused ``depth`` frames later)."""
        n = len(host_batches)
        for i in range(n + self.depth):
            if i >= self.depth:                 # retire the frame that used this slot `depth` submissions ago
                s = self.slots[i % self.depth]
                s["d2h"].synchronize()
                if on_frame is not None:
                    on_frame(i - self.depth, s["host_out"])
            if i < n:
                self._submit(self.slots[i % self.depth], host_batches[i])
        return n
