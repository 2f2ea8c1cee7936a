import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_gpu = torch.cuda.is_available()
    have_ref = os.path.isdir("/root/reference/models")
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to the scale of the reference tensor b (the 1e-5 fp32 contract)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    scale = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / scale


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)
    return load


def rows_off(a: torch.Tensor, b: torch.Tensor, tol: float) -> int:
    """Number of rows of a whose error against b exceeds tol (relative to the scale of b).  Gradients of ReLU networks
    are compared with this: a pre-activation within rounding distance of the kink flips the mask of ANY fp32 forward pass
    (the reference's included), which changes a handful of gradient rows by O(1) and leaves all others within tol."""
    a = a.detach().double().cpu().reshape(-1, a.shape[-1])
    b = b.detach().double().cpu().reshape(-1, b.shape[-1])
    scale = max(float(b.abs().max()), 1e-30)
    return int(((a - b).abs().amax(1) > tol * scale).sum())
