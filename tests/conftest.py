import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "egocentric-gaze-prediction_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.fixture(params=["precise", "precise3"])
def numeric_mode(request, monkeypatch):
    """Kernel-level tests run in both gated numeric modes (EGAZE_PRECISION): fp16-split forward + cheaper backward (default),
    and the round-1 bf16-split scheme."""
    monkeypatch.setenv("EGAZE_PRECISION", request.param)
    return request.param
