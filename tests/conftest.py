import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    import torch

    fx = torch.load(os.path.join(GOLDEN, name), weights_only=False)
    if "meta" in fx and not isinstance(fx["meta"]["transition_E"], list):
        fx["meta"] = dict(fx["meta"])
        fx["meta"]["transition_E"] = fx["meta"]["transition_E"].tolist()
    return fx


@pytest.fixture(scope="session")
def dit_small():
    return load_golden("dit_small.pt")


@pytest.fixture(scope="session")
def gin_small():
    return load_golden("gin_small.pt")
