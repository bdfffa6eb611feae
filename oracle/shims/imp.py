"""Stand-in for the ``imp`` module removed in Python 3.12; the reference's plugin loader calls
``imp.load_source(name, path)`` (lib/networks/make_network.py:8, lib/datasets/make_dataset.py)."""
import importlib.util
import sys


def load_source(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return module
