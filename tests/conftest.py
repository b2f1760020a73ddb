import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF = "/root/reference"
CKPT_CANDIDATES = [os.path.join(ROOT, "oracle", "_ref", "DeepPointMapAAAI.pth"), os.path.join(REF, "DeepPointMapAAAI.pth")]
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: imports the reference from /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(scope="session")
def checkpoint():
    for p in CKPT_CANDIDATES:
        if os.path.exists(p):
            return torch.load(p, map_location="cpu")
    pytest.skip("DeepPointMapAAAI.pth not available (oracle/_ref/ or /root/reference)")


@pytest.fixture(scope="session")
def cfg():
    from oracle import model_ref
    return model_ref.default_config()


@pytest.fixture(scope="session")
def golden_sample():
    return np.load(os.path.join(GOLDEN, "sample_pair.npz"))


@pytest.fixture(scope="session")
def golden_synth():
    return np.load(os.path.join(GOLDEN, "synthetic_8k.npz"))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (norm-wise relative error)"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
