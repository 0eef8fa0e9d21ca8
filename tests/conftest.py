import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The checker libraries are test infrastructure: build them if they are missing."""
    from softlight_b200 import build as B
    from softlight_b200 import capi as A

    if not os.path.exists(A.LIB_PATHS["port"]):
        B.build_oracle()
    yield


@pytest.fixture(scope="session")
def golden():
    frames = np.load(os.path.join(HERE, "golden", "golden_frames.npz"))
    hashes = json.load(open(os.path.join(HERE, "golden", "golden_hashes.json")))["cases"]
    return frames, hashes


@pytest.fixture(scope="session")
def golden_blits():
    return json.load(open(os.path.join(HERE, "golden", "golden_hashes.json")))["blits"]


def have_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
