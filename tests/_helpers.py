"""Shared test helpers (golden-fixture loader)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["c1_nocascade", "c2_small_cascade", "c2_headline_small", "c4_human_small", "c2_white_bkgd", "c2_no_viewdir", "c2_five_views"]
COMPOSITE_CASES = ["c5_composite_1fg", "c5_composite_2fg", "c5_composite_3fg_s6"]


def load_golden(name):
    import torch
    from enerf_b200 import config as bcfg

    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    if fx["cfg_kind"] == "composite":
        fx["cfg"] = bcfg.composite_cfg(num_fg_layers=fx["composite"], **fx["cfg_over"])
    else:
        mk = bcfg.nocascade_cfg if fx["cfg_kind"] == "nocascade" else bcfg.make_cfg
        fx["cfg"] = mk(**fx["cfg_over"])
    return fx
