import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["c1_nocascade", "c2_small_cascade", "c2_headline_small"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    import torch
    from enerf_b200 import config as bcfg

    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    mk = bcfg.nocascade_cfg if fx["cfg_kind"] == "nocascade" else bcfg.make_cfg
    fx["cfg"] = mk(**fx["cfg_over"])
    return fx


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
