"""Seeded synthetic inputs shared by the tests, bench.py and the golden-vector script
(SURVEY.md section 8d "Synthetic inputs").  No dataset or checkpoint is available offline, so the
workload is: random-init weights of the reference architecture with randomised BN statistics,
U(-1,1) source images, random pin-hole intrinsics, small random SE(3) source poses around an
identity target, near_far = [2, 6], and full-frame target rays laid out exactly as the
reference's ``lib/datasets/enerf_utils.py:60-71`` ('test' branch) produces them.
"""
import math

import numpy as np
import torch


def randomize_bn_(module_or_sd, seed=1):
    """Non-trivial BN so that folding bugs cannot hide behind the identity default."""
    g = torch.Generator().manual_seed(seed)
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else module_or_sd
    with torch.no_grad():
        for k in sorted(sd.keys()):
            v = sd[k]
            if k.endswith("running_mean"):
                v.copy_(0.1 * torch.randn(v.shape, generator=g))
            elif k.endswith("running_var"):
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
            elif ".bn." in k or _is_deconv_bn(k):
                if k.endswith(".weight"):
                    v.copy_(0.5 + torch.rand(v.shape, generator=g))
                elif k.endswith(".bias"):
                    v.copy_(0.1 * torch.randn(v.shape, generator=g))
    return module_or_sd


def _is_deconv_bn(k):
    # cost_reg_{i}.conv{7,9,11}.1.{weight,bias} are BatchNorm3d affine parameters
    parts = k.split(".")
    return len(parts) == 4 and parts[0].startswith("cost_reg_") and parts[1] in ("conv7", "conv9", "conv11") \
        and parts[2] == "1"


def _rodrigues(axis, angle):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


def full_frame_rays(tar_ext, tar_ixt, H, W, scale=1.0, row0=0, row1=None):
    """(Hs*Ws, 8) float32 rays at ``scale`` resolution: origin(3), un-normalised dir(3), u, v.
    dir = [u, v, 1] @ (K^-1)^T @ R_c2w^T   (lib/datasets/enerf_utils.py:26-32, 60-70)."""
    K = np.array(tar_ixt, dtype=np.float64).copy()
    if scale != 1.0:
        K[:2] *= scale
    Hs, Ws = int(H * scale), int(W * scale)
    row1 = Hs if row1 is None else row1
    c2w = np.linalg.inv(np.asarray(tar_ext, dtype=np.float64))
    X, Y = np.meshgrid(np.arange(Ws), np.arange(row0, row1))
    pix = np.stack([X, Y, np.ones_like(X)], axis=-1).astype(np.float64)
    dirs = pix @ (np.linalg.inv(K).T @ c2w[:3, :3].T)
    orig = np.broadcast_to(c2w[:3, 3], dirs.shape)
    rays = np.concatenate([orig, dirs, X[..., None], Y[..., None]], axis=-1)
    return rays.astype(np.float32).reshape(-1, 8)


def make_batch(H, W, S, cfg, seed=2, near_far=(2.0, 6.0), B=1):
    """The reference ``batch`` dict as CPU fp32 tensors."""
    assert B == 1, "inference batch is 1 (run.py:57-76)"
    c = cfg.enerf.cas_config
    g = torch.Generator().manual_seed(seed)
    src_inps = 2.0 * torch.rand((B, S, 3, H, W), generator=g) - 1.0
    rng = np.random.RandomState(seed + 1)
    ixts = []
    for _ in range(S + 1):
        f = rng.uniform(1.0, 1.5) * W
        cx = W / 2 * (1 + rng.uniform(-0.02, 0.02))
        cy = H / 2 * (1 + rng.uniform(-0.02, 0.02))
        ixts.append(np.array([[f, 0, cx], [0, f, cy], [0, 0, 1]], dtype=np.float64))
    rng = np.random.RandomState(seed + 2)
    exts = []
    for _ in range(S):
        R = _rodrigues(rng.normal(size=3), math.radians(rng.uniform(0, 5)))
        t = rng.uniform(-0.15, 0.15, size=3) * near_far[0]
        E = np.eye(4)
        E[:3, :3], E[:3, 3] = R, t
        exts.append(E)
    tar_ext, tar_ixt = np.eye(4), ixts[S]
    batch = {
        "src_inps": src_inps,
        "src_exts": torch.tensor(np.stack(exts)[None], dtype=torch.float32),
        "src_ixts": torch.tensor(np.stack(ixts[:S])[None], dtype=torch.float32),
        "tar_ext": torch.tensor(tar_ext[None], dtype=torch.float32),
        "tar_ixt": torch.tensor(tar_ixt[None], dtype=torch.float32),
        "near_far": torch.tensor([list(near_far)], dtype=torch.float32),
    }
    for i in range(c.num):
        batch[f"rays_{i}"] = torch.from_numpy(full_frame_rays(tar_ext, tar_ixt, H, W, c.render_scale[i]))[None]
    return batch


def make_composite_batch(H, W, S, cfg, seed=2, boxes=None):
    """``make_batch`` + the extra keys of lib/datasets/enerf_outdoor/enerf.py:183-190: one bbox (x,y,w,h;
    w,h multiples of 32 as read_tar :161-162 pads them) and one [near, far] per foreground layer, the
    background [near, far] last, and the background plates ``bg_src_inps``.  ``boxes`` overrides the
    default (x, y, w, h) per layer (w, h multiples of 32)."""
    L = int(cfg.num_fg_layers)
    batch = make_batch(H, W, S, cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 7)
    batch["bg_src_inps"] = 2.0 * torch.rand((1, S, 3, H, W), generator=g) - 1.0
    given, boxes, nfs = boxes, [], []
    for l in range(L):
        w = max(32, (W * 2 // 3) // 32 * 32) if l == 0 else 32
        h = max(32, (H // 2) // 32 * 32)
        x = min(W - w, 20 if l == 0 else 4)
        y = min(H - h, 8 if l == 0 else 24)
        boxes.append(list(given[l]) if given is not None else [x, y, w, h])
        nfs.append([2.5 + 0.5 * l, 4.5 + 1.0 * l])
    nfs.append([2.0, 6.0])
    batch["bbox"] = torch.tensor([boxes], dtype=torch.float32)
    batch["near_far"] = torch.tensor([nfs], dtype=torch.float32)
    return batch


def psnr(a, b):
    """10*log10(1/MSE), data_range 1 (lib/evaluators/enerf.py:71 uses skimage psnr, data_range=1)."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)


def make_mask_at_box(H, W, seed=9):
    """Synthetic `mask_at_box` (1,H,W) int32 like zjumocap/enerf_interactive.py:192: a centred box with
    seeded holes, so that compaction sees both long runs and isolated pixels."""
    g = torch.Generator().manual_seed(seed)
    m = torch.zeros(H, W, dtype=torch.int32)
    m[H // 5: H - H // 6, W // 4: W - W // 5] = 1
    m[torch.rand(H, W, generator=g) < 0.07] = 0
    m[0, :3] = 1
    return m[None]


# BASELINE.json configs[3] / [4] at the shapes SURVEY.md section 8d fixes (shared by tests and bench.py)
C4 = dict(H=1024, W=1024, S=4, planes=[48, 8], render_if=[False, True])          # ZJU-MoCap shape, network_human + mask_at_box
C5 = dict(H=1088, W=1920, S=6, layers=3,                                          # ENeRF-Outdoor shape (1080 padded to /32)
          boxes=[[256, 224, 384, 736], [768, 192, 416, 768], [1312, 256, 352, 704]])   # three actors, xywh multiples of 32


def c4_case(seed=2):
    """(cfg, batch) of BASELINE config 4: 1024x1024, 4 source views, 48+8 planes, masked rays."""
    from . import config as bcfg
    cfg = bcfg.make_cfg(volume_planes=list(C4["planes"]), render_if=list(C4["render_if"]))
    batch = make_batch(C4["H"], C4["W"], C4["S"], cfg, seed=seed)
    batch.pop("rays_0", None)
    batch["mask_at_box"] = make_mask_at_box(C4["H"], C4["W"])
    return cfg, batch


def c5_case(seed=2):
    """(cfg, batch) of BASELINE config 5: 1920x1088, 6 source views, 3 foreground actors + background
    (configs/enerf/enerf_outdoor/actor1.yaml: planes [32, 8] (+ background [16, 4]), samples [2, 1], no view-dir MLP)."""
    from . import config as bcfg
    cfg = bcfg.composite_cfg(num_fg_layers=C5["layers"])
    batch = make_composite_batch(C5["H"], C5["W"], C5["S"], cfg, seed=seed, boxes=C5["boxes"])
    return cfg, batch
