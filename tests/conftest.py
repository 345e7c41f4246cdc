import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build the C-ABI library once per session if it is stale or missing (nvcc
    cross-compiles without a GPU, so this also works on the CPU-only box)."""
    from audio_metrics_b200.build import build_library

    build_library()
