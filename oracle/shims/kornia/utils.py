"""Minimal stand-in for ``kornia.utils.create_meshgrid`` (un-normalised pixel grid).

Only call sites in the reference: lib/networks/enerf/utils.py:65,292, both with
``normalized_coordinates=False``.  Returns (1, H, W, 2) with [..., 0] = x in [0, W-1],
[..., 1] = y in [0, H-1] -- kornia's documented semantics."""
import torch


def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / max(width - 1, 1) - 0.5) * 2
        ys = (ys / max(height - 1, 1) - 0.5) * 2
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], dim=-1)[None]
