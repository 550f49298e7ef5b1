import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def rel_err(got, ref):
    """max |got-ref| / |ref| per entry (the north-star tolerance is <= 1e-10 on this)."""
    import torch
    return (torch.abs(got - ref) / torch.abs(ref)).max().item()


def scaled_err(got, ref):
    """max |got-ref| / max|ref| (used where entries cross zero)."""
    import torch
    return (torch.abs(got - ref).max() / torch.abs(ref).max()).item()
