"""Config access for the B200 render path.

Under the reference's ``run.py`` the global ``lib.config.cfg`` (yacs node built at import,
/root/reference/lib/config/config.py:191-201) is the source of truth and the drop-in ``Network()``
constructor reads it exactly like the reference does (lib/networks/enerf/network.py:12-22).

Stand-alone (tests, bench.py, the GPU box -- where no reference tree exists) the same keys live in a
tiny attribute-dict with the defaults of ``configs/enerf/dtu_pretrain.yaml:17-43``; ``set_cfg`` /
``make_cfg`` install one.  Only the keys the hot path reads are kept.
"""
import copy


class Node(dict):
    """Attribute-access dict (enough of yacs.CfgNode for the keys the render path reads)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return Node({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _default():
    # configs/enerf/dtu_pretrain.yaml:17-43 (+ README.md:114 overrides are applied by callers)
    return Node(
        num_fg_layers=0,
        enerf=Node(
            viewdir_agg=True,
            chunk_size=1000000,
            white_bkgd=False,
            cas_config=Node(
                num=2,
                depth_inv=[True, False],
                volume_scale=[0.125, 0.5],
                volume_planes=[64, 8],
                im_feat_scale=[0.25, 0.5],
                im_ibr_scale=[0.25, 1.0],
                render_scale=[0.25, 1.0],
                render_im_feat_level=[0, 2],
                nerf_model_feat_ch=[32, 8],
                render_if=[True, True],
                num_samples=[8, 2],
            ),
        ),
    )


_LOCAL_CFG = None


def make_cfg(**cas_overrides):
    """Fresh default cfg with ``enerf.cas_config`` keys overridden, e.g.
    ``make_cfg(volume_planes=[48, 8], render_if=[False, True])`` == the README.md:114 headline run.
    Keys ``viewdir_agg`` / ``white_bkgd`` / ``chunk_size`` go to ``cfg.enerf``."""
    cfg = _default()
    for k, v in cas_overrides.items():
        if k in ("viewdir_agg", "white_bkgd", "chunk_size"):
            cfg.enerf[k] = v
        elif k in cfg.enerf.cas_config:
            cfg.enerf.cas_config[k] = list(v) if isinstance(v, (list, tuple)) else v
        else:
            raise KeyError(f"unknown cas_config key {k!r}")
    return cfg


def nocascade_cfg(**cas_overrides):
    """configs/enerf/dtu_pretrain_nocascade.yaml:27-43 (single level at 1/4-res volume)."""
    base = dict(
        num=1, depth_inv=[True], volume_scale=[0.25], volume_planes=[48], im_feat_scale=[0.25],
        im_ibr_scale=[1.0], render_scale=[1.0], render_im_feat_level=[2], nerf_model_feat_ch=[8],
        render_if=[True], num_samples=[2],
    )
    base.update(cas_overrides)
    return make_cfg(**base)


def composite_cfg(num_fg_layers=1, **cas_overrides):
    """configs/enerf/enerf_outdoor/actor1.yaml:13-22 (network_composite: layered foreground + background)."""
    base = dict(viewdir_agg=False, volume_planes=[32, 8], num_samples=[2, 1])
    base.update(cas_overrides)
    cfg = make_cfg(**base)
    cfg.num_fg_layers = int(num_fg_layers)
    return cfg


def set_cfg(cfg):
    """Install a stand-alone cfg (ignored when the reference's lib.config is importable AND
    ``prefer_reference`` resolution finds it first; see get_cfg)."""
    global _LOCAL_CFG
    _LOCAL_CFG = cfg
    return cfg


def get_cfg():
    """Resolution order: an explicitly installed local cfg (set_cfg), else the reference's global
    ``lib.config.cfg`` when this process runs under the reference's run.py, else the defaults."""
    global _LOCAL_CFG
    if _LOCAL_CFG is not None:
        return _LOCAL_CFG
    import sys

    mod = sys.modules.get("lib.config")
    if mod is not None and hasattr(mod, "cfg"):
        return mod.cfg
    try:  # running under the reference tree (cwd on sys.path): same import the reference does
        from lib.config import cfg as ref_cfg  # type: ignore

        return ref_cfg
    except Exception:
        _LOCAL_CFG = _default()
        return _LOCAL_CFG


class LevelCfg:
    """Plain per-level snapshot of the cfg keys (taken once per forward)."""

    __slots__ = ("level", "depth_inv", "prev_depth_inv", "volume_scale", "prev_volume_scale", "planes",
                 "im_feat_scale", "im_ibr_scale", "render_scale", "im_feat_level", "feat_ch",
                 "render_if", "num_samples")

    def __init__(self, cfg, i):
        c = cfg.enerf.cas_config
        self.level = i
        self.depth_inv = bool(c.depth_inv[i])
        self.prev_depth_inv = bool(c.depth_inv[i - 1]) if i > 0 else False
        self.volume_scale = float(c.volume_scale[i])
        self.prev_volume_scale = float(c.volume_scale[i - 1]) if i > 0 else 0.0
        self.planes = int(c.volume_planes[i])
        self.im_feat_scale = float(c.im_feat_scale[i])
        self.im_ibr_scale = float(c.im_ibr_scale[i])
        self.render_scale = float(c.render_scale[i])
        self.im_feat_level = int(c.render_im_feat_level[i])
        self.feat_ch = int(c.nerf_model_feat_ch[i])
        self.render_if = bool(c.render_if[i])
        self.num_samples = int(c.num_samples[i])


def snapshot(cfg):
    n = int(cfg.enerf.cas_config.num)
    return [LevelCfg(cfg, i) for i in range(n)]
